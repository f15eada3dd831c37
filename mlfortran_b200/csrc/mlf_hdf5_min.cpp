// Minimal HDF5 writer/reader -- see mlf_hdf5_min.h.  Format follows the HDF5 File Format Specification 1.1/2.0
// for "earliest" files, which is what the reference's H5Fcreate_f/H5Dcreate_f/H5Dwrite_f sequence produces
// (src/mlf_hdf5.f90:153-166,250-264): superblock version 0, version-1 object headers, a symbol-table root group
// (local heap + v1 B-tree + SNOD leaves) and contiguous dataset layout (layout message version 3).
#include "mlf_hdf5_min.h"

#include <algorithm>
#include <cstdio>
#include <cstring>

namespace mlfh5 {
namespace {

constexpr uint64_t UNDEF = ~0ull;
constexpr int GROUP_LEAF_K = 4;       // SNOD holds up to 2*4 = 8 entries
constexpr int GROUP_INTERNAL_K = 16;  // B-tree node holds up to 32 children
constexpr uint64_t SNOD_SIZE = 8 + 2 * GROUP_LEAF_K * 40;
constexpr uint64_t BTREE_SIZE = 24 + (2 * GROUP_INTERNAL_K + 1) * 8 + 2 * GROUP_INTERNAL_K * 8;

struct Buf {
  std::vector<uint8_t> b;
  void u8(uint8_t v) { b.push_back(v); }
  void u16(uint16_t v) { for (int i = 0; i < 2; ++i) b.push_back((uint8_t)(v >> (8 * i))); }
  void u32(uint32_t v) { for (int i = 0; i < 4; ++i) b.push_back((uint8_t)(v >> (8 * i))); }
  void u64(uint64_t v) { for (int i = 0; i < 8; ++i) b.push_back((uint8_t)(v >> (8 * i))); }
  void raw(const void* p, size_t n) { const uint8_t* q = (const uint8_t*)p; b.insert(b.end(), q, q + n); }
  void pad8() { while (b.size() % 8) b.push_back(0); }
  void zeros(size_t n) { b.insert(b.end(), n, 0); }
  uint64_t size() const { return b.size(); }
  void put64(uint64_t off, uint64_t v) { for (int i = 0; i < 8; ++i) b[off + i] = (uint8_t)(v >> (8 * i)); }
};

void msg_header(Buf& o, uint16_t type, uint16_t size, uint8_t flags) {
  o.u16(type); o.u16(size); o.u8(flags); o.u8(0); o.u8(0); o.u8(0);
}

// version-1 object header holding `msgs` (already 8-byte padded message records)
void object_header(Buf& o, uint16_t nmsgs, const Buf& msgs) {
  o.u8(1); o.u8(0); o.u16(nmsgs); o.u32(1); o.u32((uint32_t)msgs.size()); o.u32(0);  // 12-byte prefix + 4 pad
  o.raw(msgs.b.data(), msgs.b.size());
}

Buf dataset_messages(const Dataset& d, uint64_t data_addr) {
  Buf m;
  const int rank = (int)d.dims.size();
  // dataspace, version 1
  msg_header(m, 0x0001, (uint16_t)(8 + 8 * rank), 0);
  m.u8(1); m.u8((uint8_t)rank); m.u8(0); m.u8(0); m.u32(0);
  for (uint64_t v : d.dims) m.u64(v);
  m.pad8();
  // datatype, version 1
  if (d.type == I64) {
    msg_header(m, 0x0003, 16, 1 /*constant*/);
    m.u8(0x10 | 0); m.u8(0x08); m.u8(0); m.u8(0); m.u32(8);  // fixed-point, little-endian, signed
    m.u16(0); m.u16(64);
    m.pad8();
  } else {
    msg_header(m, 0x0003, 24, 1);
    m.u8(0x10 | 1); m.u8(0x20); m.u8(0x3F); m.u8(0); m.u32(8);  // IEEE double LE: msb-implied mantissa, sign bit 63
    m.u16(0); m.u16(64); m.u8(52); m.u8(11); m.u8(0); m.u8(52); m.u32(1023);
    m.pad8();
  }
  // fill value, version 2: late allocation, fill time "if set", default (zero-length) value
  msg_header(m, 0x0005, 8, 1);
  m.u8(2); m.u8(2); m.u8(2); m.u8(1); m.u32(0);
  // data layout, version 3, contiguous
  msg_header(m, 0x0008, 24, 0);
  m.u8(3); m.u8(1); m.u64(data_addr); m.u64(d.count() * 8);
  m.pad8();
  return m;
}

}  // namespace

int write_file(const std::string& path, const std::vector<Dataset>& dsin) {
  std::vector<const Dataset*> ds;
  for (const auto& d : dsin) ds.push_back(&d);
  std::sort(ds.begin(), ds.end(), [](const Dataset* a, const Dataset* b) { return a->name < b->name; });  // strcmp order
  for (size_t i = 1; i < ds.size(); ++i)
    if (ds[i]->name == ds[i - 1]->name) return -6;
  const size_t n = ds.size();
  const size_t nsnod = std::max<size_t>(1, (n + 2 * GROUP_LEAF_K - 1) / (2 * GROUP_LEAF_K));
  if (nsnod > 2 * GROUP_INTERNAL_K) return -7;

  // ---- local heap data segment: "" at offset 0, then the names, then one free block --------------------------
  Buf heap;
  heap.zeros(8);
  std::vector<uint64_t> name_off(n);
  for (size_t i = 0; i < n; ++i) {
    name_off[i] = heap.size();
    heap.raw(ds[i]->name.c_str(), ds[i]->name.size() + 1);
    heap.pad8();
  }
  const uint64_t free_off = heap.size();
  const uint64_t free_len = 32;
  heap.u64(1);  // H5HL_FREE_NULL: last free block
  heap.u64(free_len);
  heap.zeros(free_len - 16);

  // ---- layout -----------------------------------------------------------------------------------------------
  uint64_t at = 96;  // superblock v0 with 8-byte offsets/lengths
  const uint64_t root_oh = at;
  Buf rootmsgs;
  msg_header(rootmsgs, 0x0011, 16, 0);
  const uint64_t root_stab_patch = rootmsgs.size();
  rootmsgs.u64(0); rootmsgs.u64(0);
  at += 16 + rootmsgs.size();
  const uint64_t btree_addr = at; at += BTREE_SIZE;
  const uint64_t heap_addr = at; at += 32;
  const uint64_t heapdata_addr = at; at += heap.size();
  std::vector<uint64_t> snod_addr(nsnod);
  for (size_t s = 0; s < nsnod; ++s) { snod_addr[s] = at; at += SNOD_SIZE; }
  std::vector<uint64_t> oh_addr(n), data_addr(n);
  std::vector<Buf> ohs(n);
  for (size_t i = 0; i < n; ++i) {  // header sizes do not depend on the address values
    Buf m = dataset_messages(*ds[i], 0);
    oh_addr[i] = at;
    at += 16 + m.size();
  }
  for (size_t i = 0; i < n; ++i) {
    data_addr[i] = at;
    at += (ds[i]->count() * 8 + 7) / 8 * 8;
    if (ds[i]->data.size() != ds[i]->count() * 8) return -7;
  }
  const uint64_t eof = at;
  rootmsgs.put64(root_stab_patch, btree_addr);
  rootmsgs.put64(root_stab_patch + 8, heap_addr);

  // ---- emit ---------------------------------------------------------------------------------------------------
  Buf f;
  const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
  f.raw(sig, 8);
  f.u8(0); f.u8(0); f.u8(0); f.u8(0); f.u8(0); f.u8(8); f.u8(8); f.u8(0);
  f.u16(GROUP_LEAF_K); f.u16(GROUP_INTERNAL_K); f.u32(0);
  f.u64(0); f.u64(UNDEF); f.u64(eof); f.u64(UNDEF);
  f.u64(0); f.u64(root_oh); f.u32(1); f.u32(0); f.u64(btree_addr); f.u64(heap_addr);  // root symbol-table entry
  if (f.size() != 96) return -7;
  object_header(f, 1, rootmsgs);
  // B-tree leaf node over the SNODs
  f.raw("TREE", 4); f.u8(0); f.u8(0); f.u16((uint16_t)(n ? nsnod : 0)); f.u64(UNDEF); f.u64(UNDEF);
  {
    Buf kc;
    kc.u64(0);  // key 0: the empty string
    for (size_t s = 0; s < nsnod && n; ++s) {
      kc.u64(snod_addr[s]);
      const size_t last = std::min(n, (s + 1) * 2 * GROUP_LEAF_K) - 1;
      kc.u64(name_off[last]);  // largest name in child s
    }
    kc.zeros(BTREE_SIZE - 24 - kc.size());
    f.raw(kc.b.data(), kc.b.size());
  }
  // local heap
  f.raw("HEAP", 4); f.u8(0); f.u8(0); f.u8(0); f.u8(0); f.u64(heap.size()); f.u64(free_off); f.u64(heapdata_addr);
  f.raw(heap.b.data(), heap.b.size());
  // symbol-table nodes
  for (size_t s = 0; s < nsnod; ++s) {
    const size_t lo = s * 2 * GROUP_LEAF_K, hi = std::min(n, lo + 2 * GROUP_LEAF_K);
    f.raw("SNOD", 4); f.u8(1); f.u8(0); f.u16((uint16_t)(hi > lo ? hi - lo : 0));
    for (size_t i = lo; i < hi; ++i) { f.u64(name_off[i]); f.u64(oh_addr[i]); f.u32(0); f.u32(0); f.zeros(16); }
    f.zeros((2 * GROUP_LEAF_K - (hi > lo ? hi - lo : 0)) * 40);
  }
  for (size_t i = 0; i < n; ++i) {
    if (f.size() != oh_addr[i]) return -7;
    object_header(f, 4, dataset_messages(*ds[i], data_addr[i]));
  }
  for (size_t i = 0; i < n; ++i) {
    if (f.size() != data_addr[i]) return -7;
    f.raw(ds[i]->data.data(), ds[i]->data.size());
    f.pad8();
  }
  if (f.size() != eof) return -7;
  FILE* fp = fopen(path.c_str(), "wb");
  if (!fp) return -6;
  const size_t w = fwrite(f.b.data(), 1, f.b.size(), fp);
  const int c = fclose(fp);
  return (w == f.b.size() && c == 0) ? 0 : -6;
}

// ---------------------------------------------------------------------------------------------------------------
namespace {

struct Reader {
  std::vector<uint8_t> f;
  bool ok(uint64_t off, uint64_t n) const { return off <= f.size() && n <= f.size() - off; }
  uint64_t u(uint64_t off, int n) const {
    uint64_t v = 0;
    for (int i = 0; i < n; ++i) v |= (uint64_t)f[off + i] << (8 * i);
    return v;
  }
};

struct Msg { uint16_t type; uint64_t off; uint16_t size; };

// collects the messages of a version-1 object header, following continuation messages (0x0010)
bool read_messages(const Reader& r, uint64_t oh, std::vector<Msg>& out) {
  if (!r.ok(oh, 16) || r.f[oh] != 1) return false;
  const unsigned nmsgs = (unsigned)r.u(oh + 2, 2);
  std::vector<std::pair<uint64_t, uint64_t>> chunks{{oh + 16, r.u(oh + 8, 4)}};
  for (size_t c = 0; c < chunks.size() && out.size() < nmsgs; ++c) {
    uint64_t p = chunks[c].first;
    const uint64_t end = p + chunks[c].second;
    if (!r.ok(p, chunks[c].second)) return false;
    while (p + 8 <= end && out.size() < nmsgs) {
      Msg m{(uint16_t)r.u(p, 2), p + 8, (uint16_t)r.u(p + 2, 2)};
      if (m.off + m.size > end) return false;
      if (m.type == 0x0010 && m.size >= 16) chunks.push_back({r.u(m.off, 8), r.u(m.off + 8, 8)});
      out.push_back(m);
      p = m.off + m.size;
    }
  }
  return true;
}

bool read_dataset(const Reader& r, uint64_t oh, Dataset& d) {
  std::vector<Msg> msgs;
  if (!read_messages(r, oh, msgs)) return false;
  bool have_space = false, have_type = false, have_layout = false;
  uint64_t addr = UNDEF, bytes = 0;
  for (const Msg& m : msgs) {
    if (m.type == 0x0001) {  // dataspace
      const int ver = r.f[m.off], rank = r.f[m.off + 1];
      uint64_t p = m.off + (ver == 1 ? 8 : 4);
      if (ver != 1 && ver != 2) return false;
      d.dims.clear();
      for (int i = 0; i < rank; ++i) d.dims.push_back(r.u(p + 8 * i, 8));
      have_space = true;
    } else if (m.type == 0x0003) {  // datatype
      const int cls = r.f[m.off] & 0x0F;
      const uint32_t size = (uint32_t)r.u(m.off + 4, 4);
      const bool big_endian = r.f[m.off + 1] & 1;
      if (size != 8 || big_endian) return false;
      if (cls == 0) d.type = I64;
      else if (cls == 1) d.type = F64;
      else return false;
      have_type = true;
    } else if (m.type == 0x0008) {  // layout
      const int ver = r.f[m.off];
      if (ver == 3) {
        if (r.f[m.off + 1] != 1) return false;  // contiguous only
        addr = r.u(m.off + 2, 8);
        bytes = r.u(m.off + 10, 8);
      } else if (ver == 1 || ver == 2) {
        const int rank = r.f[m.off + 1];
        if (r.f[m.off + 2] != 1) return false;
        addr = r.u(m.off + 8, 8);
        bytes = 0;
        (void)rank;
      } else return false;
      have_layout = true;
    }
  }
  if (!have_space || !have_type || !have_layout) return false;
  const uint64_t need = d.count() * 8;
  if (bytes == 0) bytes = need;
  if (bytes < need) return false;
  d.data.assign(need, 0);
  if (addr != UNDEF) {  // never-written datasets have no storage and read as the (zero) fill value
    if (!r.ok(addr, need)) return false;
    std::memcpy(d.data.data(), r.f.data() + addr, need);
  }
  return true;
}

bool walk_btree(const Reader& r, uint64_t node, uint64_t heapdata, std::vector<Dataset>& out, int depth) {
  if (depth > 8 || !r.ok(node, 24) || std::memcmp(&r.f[node], "TREE", 4) != 0 || r.f[node + 4] != 0) return false;
  const int level = r.f[node + 5];
  const unsigned used = (unsigned)r.u(node + 6, 2);
  uint64_t p = node + 24;
  for (unsigned i = 0; i < used; ++i) {
    if (!r.ok(p, 24)) return false;
    const uint64_t child = r.u(p + 8, 8);
    p += 16;
    if (level > 0) {
      if (!walk_btree(r, child, heapdata, out, depth + 1)) return false;
      continue;
    }
    if (!r.ok(child, 8) || std::memcmp(&r.f[child], "SNOD", 4) != 0) return false;
    const unsigned ns = (unsigned)r.u(child + 6, 2);
    for (unsigned s = 0; s < ns; ++s) {
      const uint64_t e = child + 8 + 40ull * s;
      if (!r.ok(e, 40)) return false;
      const uint64_t noff = heapdata + r.u(e, 8);
      if (!r.ok(noff, 1)) return false;
      Dataset d;
      for (uint64_t q = noff; q < r.f.size() && r.f[q]; ++q) d.name.push_back((char)r.f[q]);
      if (read_dataset(r, r.u(e + 8, 8), d)) out.push_back(std::move(d));  // groups / unsupported types are skipped
    }
  }
  return true;
}

}  // namespace

int read_file(const std::string& path, std::vector<Dataset>& out) {
  FILE* fp = fopen(path.c_str(), "rb");
  if (!fp) return -5;
  Reader r;
  fseek(fp, 0, SEEK_END);
  const long sz = ftell(fp);
  fseek(fp, 0, SEEK_SET);
  r.f.resize(sz > 0 ? (size_t)sz : 0);
  const size_t got = r.f.empty() ? 0 : fread(r.f.data(), 1, r.f.size(), fp);
  fclose(fp);
  if (got != r.f.size() || r.f.size() < 96) return -6;
  const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
  if (std::memcmp(r.f.data(), sig, 8) != 0) return -6;
  const int ver = r.f[8];
  if (ver > 1 || r.f[13] != 8 || r.f[14] != 8) return -6;  // superblock v0/v1 with 8-byte offsets and lengths
  const uint64_t ste = (ver == 0 ? 56 : 60);                // root symbol-table entry
  const uint64_t root_oh = r.u(ste + 8, 8);
  uint64_t btree = UNDEF, heap = UNDEF;
  if (r.u(ste + 16, 4) == 1) { btree = r.u(ste + 24, 8); heap = r.u(ste + 32, 8); }
  else {
    std::vector<Msg> msgs;
    if (!read_messages(r, root_oh, msgs)) return -6;
    for (const Msg& m : msgs)
      if (m.type == 0x0011) { btree = r.u(m.off, 8); heap = r.u(m.off + 8, 8); }
  }
  if (btree == UNDEF || heap == UNDEF || !r.ok(heap, 32) || std::memcmp(&r.f[heap], "HEAP", 4) != 0) return -6;
  const uint64_t heapdata = r.u(heap + 24, 8);
  out.clear();
  return walk_btree(r, btree, heapdata, out, 0) ? 0 : -6;
}

}  // namespace mlfh5
