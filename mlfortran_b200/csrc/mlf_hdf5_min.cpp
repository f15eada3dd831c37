// Minimal HDF5 writer/reader -- see mlf_hdf5_min.h.  Format follows the HDF5 File Format Specification 1.1/2.0
// for "earliest" files, which is what the reference's H5Fcreate_f/H5Gcreate_f/H5Dcreate_f/H5Dwrite_f sequence produces
// (src/mlf_hdf5.f90:106-131,153-166,250-264): superblock version 0, version-1 object headers, symbol-table groups
// (local heap + v1 B-tree + SNOD leaves) and contiguous dataset layout (layout message version 3).
#include "mlf_hdf5_min.h"

#include <algorithm>
#include <climits>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <set>

#include <unistd.h>

namespace mlfh5 {
namespace {

constexpr uint64_t UNDEF = ~0ull;
constexpr int GROUP_LEAF_K = 4;       // SNOD holds up to 2*4 = 8 entries
constexpr int GROUP_INTERNAL_K = 16;  // B-tree node holds up to 32 children
constexpr uint64_t SNOD_SIZE = 8 + 2 * GROUP_LEAF_K * 40;
constexpr uint64_t BTREE_SIZE = 24 + (2 * GROUP_INTERNAL_K + 1) * 8 + 2 * GROUP_INTERNAL_K * 8;
constexpr uint64_t GROUP_OH_SIZE = 16 + 8 + 16;   // object header prefix + one symbol-table message
constexpr int MAX_RANK = 15;                      // MLF_MAXRANK

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

// ---- writer: group tree -------------------------------------------------------------------------------------------
struct Group {
  std::map<std::string, std::unique_ptr<Group>> sub;   // std::map orders by operator< on std::string == strcmp order
  std::map<std::string, const Dataset*> ds;
  // entries of this group in name order: (name, is_group)
  std::vector<std::pair<std::string, bool>> entries;
  std::vector<uint64_t> name_off;
  Buf heap;
  uint64_t free_off = 0;
  uint64_t oh = 0, btree = 0, heap_addr = 0, heapdata = 0;
  std::vector<uint64_t> snod;
  std::map<std::string, uint64_t> ds_oh, ds_data;
};

bool valid_component(const std::string& s) { return !s.empty() && s != "." && s != ".." && s.find('\0') == std::string::npos; }

Group* descend(Group& root, const std::string& path, std::string* leaf) {
  Group* g = &root;
  size_t p = 0;
  while (true) {
    const size_t q = path.find('/', p);
    const std::string comp = path.substr(p, q == std::string::npos ? std::string::npos : q - p);
    if (!valid_component(comp)) return nullptr;
    if (q == std::string::npos) {
      if (leaf) { *leaf = comp; return g; }
      // the whole path names a group
      if (g->ds.count(comp)) return nullptr;
      auto& s = g->sub[comp];
      if (!s) s.reset(new Group());
      return s.get();
    }
    if (g->ds.count(comp)) return nullptr;  // a dataset where a group is needed
    auto& s = g->sub[comp];
    if (!s) s.reset(new Group());
    g = s.get();
    p = q + 1;
  }
}

int prepare(Group& g) {
  g.entries.clear();
  for (auto& kv : g.sub) g.entries.push_back({kv.first, true});
  for (auto& kv : g.ds) {
    if (g.sub.count(kv.first)) return -6;
    g.entries.push_back({kv.first, false});
  }
  std::sort(g.entries.begin(), g.entries.end());
  const size_t n = g.entries.size();
  if ((n + 2 * GROUP_LEAF_K - 1) / (2 * GROUP_LEAF_K) > 2 * GROUP_INTERNAL_K) return -7;  // one B-tree node per group
  g.heap = Buf();
  g.heap.zeros(8);  // "" at offset 0
  g.name_off.resize(n);
  for (size_t i = 0; i < n; ++i) {
    g.name_off[i] = g.heap.size();
    g.heap.raw(g.entries[i].first.c_str(), g.entries[i].first.size() + 1);
    g.heap.pad8();
  }
  g.free_off = g.heap.size();
  g.heap.u64(1);   // H5HL_FREE_NULL: last free block
  g.heap.u64(32);
  g.heap.zeros(16);
  for (auto& kv : g.sub)
    if (int r = prepare(*kv.second)) return r;
  return 0;
}

// metadata of a group and everything below it, depth first; dataset payloads are placed afterwards
void layout_meta(Group& g, uint64_t& at) {
  g.oh = at; at += GROUP_OH_SIZE;
  g.btree = at; at += BTREE_SIZE;
  g.heap_addr = at; at += 32;
  g.heapdata = at; at += g.heap.size();
  const size_t n = g.entries.size();
  const size_t nsnod = std::max<size_t>(1, (n + 2 * GROUP_LEAF_K - 1) / (2 * GROUP_LEAF_K));
  g.snod.resize(nsnod);
  for (size_t s = 0; s < nsnod; ++s) { g.snod[s] = at; at += SNOD_SIZE; }
  for (auto& kv : g.ds) {
    g.ds_oh[kv.first] = at;
    at += 16 + dataset_messages(*kv.second, 0).size();   // header sizes do not depend on the address values
  }
  for (auto& kv : g.sub) layout_meta(*kv.second, at);
}
void layout_data(Group& g, uint64_t& at) {
  for (auto& kv : g.ds) {
    g.ds_data[kv.first] = at;
    at += (kv.second->count() * 8 + 7) / 8 * 8;
  }
  for (auto& kv : g.sub) layout_data(*kv.second, at);
}

bool emit_meta(const Group& g, Buf& f) {
  if (f.size() != g.oh) return false;
  Buf gm;
  msg_header(gm, 0x0011, 16, 0);
  gm.u64(g.btree); gm.u64(g.heap_addr);
  object_header(f, 1, gm);
  const size_t n = g.entries.size(), nsnod = g.snod.size();
  // B-tree leaf node over the SNODs
  f.raw("TREE", 4); f.u8(0); f.u8(0); f.u16((uint16_t)(n ? nsnod : 0)); f.u64(UNDEF); f.u64(UNDEF);
  {
    Buf kc;
    kc.u64(0);  // key 0: the empty string
    for (size_t s = 0; s < nsnod && n; ++s) {
      kc.u64(g.snod[s]);
      const size_t last = std::min(n, (s + 1) * 2 * GROUP_LEAF_K) - 1;
      kc.u64(g.name_off[last]);  // largest name in child s
    }
    kc.zeros(BTREE_SIZE - 24 - kc.size());
    f.raw(kc.b.data(), kc.b.size());
  }
  f.raw("HEAP", 4); f.u8(0); f.u8(0); f.u8(0); f.u8(0); f.u64(g.heap.size()); f.u64(g.free_off); f.u64(g.heapdata);
  f.raw(g.heap.b.data(), g.heap.b.size());
  for (size_t s = 0; s < nsnod; ++s) {
    const size_t lo = s * 2 * GROUP_LEAF_K, hi = std::min(n, lo + 2 * GROUP_LEAF_K);
    f.raw("SNOD", 4); f.u8(1); f.u8(0); f.u16((uint16_t)(hi > lo ? hi - lo : 0));
    for (size_t i = lo; i < hi; ++i) {
      const std::string& nm = g.entries[i].first;
      f.u64(g.name_off[i]);
      if (g.entries[i].second) {   // sub-group: cached symbol-table information, as libhdf5 writes it
        const Group& c = *g.sub.at(nm);
        f.u64(c.oh); f.u32(1); f.u32(0); f.u64(c.btree); f.u64(c.heap_addr);
      } else {
        f.u64(g.ds_oh.at(nm)); f.u32(0); f.u32(0); f.zeros(16);
      }
    }
    f.zeros((2 * GROUP_LEAF_K - (hi > lo ? hi - lo : 0)) * 40);
  }
  for (auto& kv : g.ds) {
    if (f.size() != g.ds_oh.at(kv.first)) return false;
    object_header(f, 4, dataset_messages(*kv.second, g.ds_data.at(kv.first)));
  }
  for (auto& kv : g.sub)
    if (!emit_meta(*kv.second, f)) return false;
  return true;
}
bool emit_data(const Group& g, Buf& f) {
  for (auto& kv : g.ds) {
    if (f.size() != g.ds_data.at(kv.first)) return false;
    f.raw(kv.second->data.data(), kv.second->data.size());
    f.pad8();
  }
  for (auto& kv : g.sub)
    if (!emit_data(*kv.second, f)) return false;
  return true;
}

}  // namespace

int write_file(const std::string& path, const Contents& c) {
  Group root;
  for (const std::string& gp : c.groups)
    if (!descend(root, gp, nullptr)) return -6;
  for (const Dataset& d : c.ds) {
    if (d.dims.empty() || d.dims.size() > (size_t)MAX_RANK || d.data.size() != d.count() * 8) return -7;
    std::string leaf;
    Group* g = descend(root, d.name, &leaf);
    if (!g || g->ds.count(leaf) || g->sub.count(leaf)) return -6;   // bad path or duplicate name
    g->ds[leaf] = &d;
  }
  if (int r = prepare(root)) return r;
  uint64_t at = 96;  // superblock v0 with 8-byte offsets/lengths
  layout_meta(root, at);
  layout_data(root, at);
  const uint64_t eof = at;

  Buf f;
  f.b.reserve((size_t)eof);
  const uint8_t sig[8] = {0x89, 'H', 'D', 'F', '\r', '\n', 0x1a, '\n'};
  f.raw(sig, 8);
  f.u8(0); f.u8(0); f.u8(0); f.u8(0); f.u8(0); f.u8(8); f.u8(8); f.u8(0);
  f.u16(GROUP_LEAF_K); f.u16(GROUP_INTERNAL_K); f.u32(0);
  f.u64(0); f.u64(UNDEF); f.u64(eof); f.u64(UNDEF);
  f.u64(0); f.u64(root.oh); f.u32(1); f.u32(0); f.u64(root.btree); f.u64(root.heap_addr);  // root symbol-table entry
  if (f.size() != 96) return -7;
  if (!emit_meta(root, f) || !emit_data(root, f) || f.size() != eof) return -7;

  // temporary beside the target, then rename(2): readers and crashes see the old file or the new one, never a torso
  const std::string tmp = path + ".tmp." + std::to_string((long)getpid());
  FILE* fp = fopen(tmp.c_str(), "wb");
  if (!fp) return -6;
  const size_t w = fwrite(f.b.data(), 1, f.b.size(), fp);
  const int fl = fflush(fp);
  const int c2 = fclose(fp);
  if (w != f.b.size() || fl != 0 || c2 != 0 || rename(tmp.c_str(), path.c_str()) != 0) {
    remove(tmp.c_str());
    return -6;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------
namespace {

struct Reader {
  std::vector<uint8_t> f;
  uint64_t base = 0;   // address of the superblock: every file address is relative to it
  bool ok(uint64_t off, uint64_t n) const { return off <= f.size() && n <= f.size() - off; }
  // bounds-checked little-endian read of n <= 8 bytes at absolute offset off
  bool get(uint64_t off, int n, uint64_t& v) const {
    if (!ok(off, (uint64_t)n)) return false;
    v = 0;
    for (int i = 0; i < n; ++i) v |= (uint64_t)f[off + i] << (8 * i);
    return true;
  }
  // file address -> absolute offset (false on overflow / undefined address)
  bool abs(uint64_t addr, uint64_t& off) const {
    if (addr == UNDEF || addr > f.size()) return false;
    off = base + addr;
    return off <= f.size();
  }
};

struct Msg { uint16_t type; uint64_t off; uint16_t size; uint8_t flags; };

// collects the messages of a version-1 object header, following continuation messages (0x0010)
bool read_messages(const Reader& r, uint64_t oh, std::vector<Msg>& out) {
  uint64_t ver, nmsgs, hsize;
  if (!r.get(oh, 1, ver) || ver != 1 || !r.get(oh + 2, 2, nmsgs) || !r.get(oh + 8, 4, hsize)) return false;
  std::vector<std::pair<uint64_t, uint64_t>> chunks{{oh + 16, hsize}};
  for (size_t c = 0; c < chunks.size() && out.size() < nmsgs; ++c) {
    if (c > 64) return false;
    uint64_t p = chunks[c].first;
    if (!r.ok(p, chunks[c].second)) return false;
    const uint64_t end = p + chunks[c].second;
    while (p + 8 <= end && out.size() < nmsgs) {
      uint64_t type, size, flags;
      if (!r.get(p, 2, type) || !r.get(p + 2, 2, size) || !r.get(p + 4, 1, flags)) return false;
      Msg m{(uint16_t)type, p + 8, (uint16_t)size, (uint8_t)flags};
      if (m.off + m.size > end) return false;
      if (m.type == 0x0010) {
        uint64_t caddr, clen, coff;
        if (m.size < 16 || !r.get(m.off, 8, caddr) || !r.get(m.off + 8, 8, clen) || !r.abs(caddr, coff)) return false;
        chunks.push_back({coff, clen});
      }
      out.push_back(m);
      p = m.off + m.size;
    }
  }
  return true;
}

enum ObjKind { OBJ_BAD, OBJ_GROUP, OBJ_DATASET, OBJ_UNSUPPORTED };

// Classifies the object behind header `oh`; fills `d` for a supported dataset, (btree, heap) for a group.
ObjKind read_object(const Reader& r, uint64_t oh, Dataset& d, uint64_t& btree, uint64_t& heap, bool& has_attr) {
  std::vector<Msg> msgs;
  if (!read_messages(r, oh, msgs)) return OBJ_BAD;
  bool have_space = false, have_type = false, have_layout = false, unsupported = false;
  uint64_t addr = UNDEF, bytes = 0;
  has_attr = false;
  for (const Msg& m : msgs) {
    if (m.flags & 2) { unsupported = true; continue; }   // shared message: stored elsewhere
    if (m.type == 0x0011) {  // symbol table: this is a group
      if (m.size < 16 || !r.get(m.off, 8, btree) || !r.get(m.off + 8, 8, heap)) return OBJ_BAD;
      return OBJ_GROUP;
    } else if (m.type == 0x0001) {  // dataspace
      uint64_t ver, rank;
      if (m.size < 4 || !r.get(m.off, 1, ver) || !r.get(m.off + 1, 1, rank)) return OBJ_BAD;
      if (ver != 1 && ver != 2) { unsupported = true; continue; }
      if (rank > (uint64_t)MAX_RANK) return OBJ_BAD;
      const uint64_t p = m.off + (ver == 1 ? 8 : 4);
      if (p + 8 * rank > m.off + m.size) return OBJ_BAD;
      d.dims.clear();
      for (uint64_t i = 0; i < rank; ++i) {
        uint64_t v;
        if (!r.get(p + 8 * i, 8, v)) return OBJ_BAD;
        d.dims.push_back(v);
      }
      if (rank == 0) unsupported = true;  // scalar / null dataspaces are not part of the checkpoint layout
      have_space = true;
    } else if (m.type == 0x0003) {  // datatype
      uint64_t cv, b0, b1, size;
      if (m.size < 8 || !r.get(m.off, 1, cv) || !r.get(m.off + 1, 1, b0) || !r.get(m.off + 2, 1, b1) || !r.get(m.off + 4, 4, size)) return OBJ_BAD;
      const int cls = (int)(cv & 0x0F);
      const bool big_endian = b0 & 1;
      if (size != 8 || big_endian) { unsupported = true; continue; }
      if (cls == 0 && (b0 & 0x08)) d.type = I64;            // signed fixed-point
      else if (cls == 1) {
        uint64_t esize, msize;
        if (m.size < 20 || !r.get(m.off + 13, 1, esize) || !r.get(m.off + 15, 1, msize)) return OBJ_BAD;
        if (esize != 11 || msize != 52) { unsupported = true; continue; }
        d.type = F64;
      } else { unsupported = true; continue; }
      have_type = true;
    } else if (m.type == 0x0008) {  // layout
      uint64_t ver, cls;
      if (m.size < 2 || !r.get(m.off, 1, ver)) return OBJ_BAD;
      if (ver == 3) {
        if (!r.get(m.off + 1, 1, cls)) return OBJ_BAD;
        if (cls != 1) { unsupported = true; continue; }  // contiguous only
        if (m.size < 18 || !r.get(m.off + 2, 8, addr) || !r.get(m.off + 10, 8, bytes)) return OBJ_BAD;
      } else if (ver == 1 || ver == 2) {
        if (m.size < 16 || !r.get(m.off + 2, 1, cls)) return OBJ_BAD;
        if (cls != 1) { unsupported = true; continue; }
        if (!r.get(m.off + 8, 8, addr)) return OBJ_BAD;
        bytes = 0;
      } else { unsupported = true; continue; }
      have_layout = true;
    } else if (m.type == 0x000C) {
      has_attr = true;
    } else if (m.type == 0x000B || m.type == 0x0007) {  // filter pipeline, external data files
      unsupported = true;
    }
  }
  if (unsupported || !have_space || !have_type || !have_layout) return OBJ_UNSUPPORTED;
  uint64_t need = 8;
  for (uint64_t v : d.dims) {
    if (v != 0 && need > (uint64_t)INT_MAX * 8ull / v) return OBJ_BAD;   // count*8 must not wrap (and dims must fit an int)
    if (v > (uint64_t)INT_MAX) return OBJ_BAD;
    need *= v;
  }
  if (bytes == 0) bytes = need;
  if (bytes < need) return OBJ_BAD;
  if (need > r.f.size()) return OBJ_BAD;
  d.data.assign((size_t)need, 0);
  if (addr != UNDEF) {  // never-written datasets have no storage and read as the (zero) fill value
    uint64_t off;
    if (!r.abs(addr, off) || !r.ok(off, need)) return OBJ_BAD;
    std::memcpy(d.data.data(), r.f.data() + off, (size_t)need);
  }
  return OBJ_DATASET;
}

bool walk_group(const Reader& r, uint64_t btree, uint64_t heap, const std::string& prefix, Contents& out, std::set<uint64_t>& seen, int depth);

bool walk_btree(const Reader& r, uint64_t node, uint64_t heapdata, uint64_t heapsize, const std::string& prefix, Contents& out,
                std::set<uint64_t>& seen, int depth, int level_guard) {
  uint64_t noff;
  if (level_guard > 8 || !r.abs(node, noff) || !r.ok(noff, 24) || std::memcmp(&r.f[noff], "TREE", 4) != 0 || r.f[noff + 4] != 0) return false;
  const int level = r.f[noff + 5];
  uint64_t used;
  if (!r.get(noff + 6, 2, used) || used > 2 * 1024) return false;
  uint64_t p = noff + 24;
  for (uint64_t i = 0; i < used; ++i) {
    uint64_t child;
    if (!r.get(p + 8, 8, child)) return false;
    p += 16;
    if (level > 0) {
      if (!walk_btree(r, child, heapdata, heapsize, prefix, out, seen, depth, level_guard + 1)) return false;
      continue;
    }
    uint64_t coff, ns;
    if (!r.abs(child, coff) || !r.ok(coff, 8) || std::memcmp(&r.f[coff], "SNOD", 4) != 0 || !r.get(coff + 6, 2, ns) || ns > 4096) return false;
    for (uint64_t s = 0; s < ns; ++s) {
      const uint64_t e = coff + 8 + 40ull * s;
      uint64_t nameo, oh, ctype;
      if (!r.get(e, 8, nameo) || !r.get(e + 8, 8, oh) || !r.get(e + 16, 4, ctype)) return false;
      if (nameo >= heapsize) return false;
      std::string name;
      bool terminated = false;
      for (uint64_t q = heapdata + nameo; q < heapdata + heapsize && r.ok(q, 1); ++q) {
        if (!r.f[q]) { terminated = true; break; }
        name.push_back((char)r.f[q]);
      }
      if (!terminated || name.empty()) return false;
      const std::string path = prefix + name;
      if (ctype == 2) { out.lossy.push_back(path + " (symbolic link)"); continue; }
      uint64_t ohoff;
      if (!r.abs(oh, ohoff)) return false;
      Dataset d;
      d.name = path;
      uint64_t gb = UNDEF, gh = UNDEF;
      bool has_attr = false;
      switch (read_object(r, ohoff, d, gb, gh, has_attr)) {
        case OBJ_BAD: return false;
        case OBJ_GROUP:
          if (!seen.insert(oh).second) { out.lossy.push_back(path + " (hard link to a group already visited)"); break; }
          out.groups.push_back(path);
          if (!walk_group(r, gb, gh, path + "/", out, seen, depth + 1)) return false;
          break;
        case OBJ_DATASET:
          if (has_attr) out.lossy.push_back(path + " (attributes)");
          out.ds.push_back(std::move(d));
          break;
        case OBJ_UNSUPPORTED: out.lossy.push_back(path + " (type / layout outside int64 | float64, contiguous)"); break;
      }
    }
  }
  return true;
}

bool walk_group(const Reader& r, uint64_t btree, uint64_t heap, const std::string& prefix, Contents& out, std::set<uint64_t>& seen, int depth) {
  if (depth > 32) return false;
  uint64_t hoff, hsize, hdata, hdoff;
  if (!r.abs(heap, hoff) || !r.ok(hoff, 32) || std::memcmp(&r.f[hoff], "HEAP", 4) != 0) return false;
  if (!r.get(hoff + 8, 8, hsize) || !r.get(hoff + 24, 8, hdata) || !r.abs(hdata, hdoff) || !r.ok(hdoff, hsize)) return false;
  return walk_btree(r, btree, hdoff, hsize, prefix, out, seen, depth, 0);
}

}  // namespace

int read_file(const std::string& path, Contents& out) {
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
  // the superblock sits at 0 or, behind a user block, at 512, 1024, 2048, ...
  uint64_t sb = UNDEF;
  for (uint64_t cand = 0; cand + 96 <= r.f.size(); cand = cand ? cand * 2 : 512)
    if (std::memcmp(r.f.data() + cand, sig, 8) == 0) { sb = cand; break; }
  if (sb == UNDEF) return -6;
  const int ver = r.f[sb + 8];
  if (ver > 1 || r.f[sb + 13] != 8 || r.f[sb + 14] != 8) return -6;  // superblock v0/v1 with 8-byte offsets and lengths
  const uint64_t fixed = sb + (ver == 0 ? 24 : 28);                  // base address, free-space, EOF, driver-info addresses
  uint64_t base_addr;
  if (!r.get(fixed, 8, base_addr)) return -6;
  r.base = base_addr == 0 ? sb : base_addr;   // libhdf5 sets the base address to the user-block size
  if (r.base != sb) return -6;
  const uint64_t ste = fixed + 32;            // root symbol-table entry
  uint64_t root_oh, ctype, btree = UNDEF, heap = UNDEF;
  if (!r.get(ste + 8, 8, root_oh) || !r.get(ste + 16, 4, ctype)) return -6;
  out = Contents();
  if (sb != 0) out.lossy.push_back("(user block of " + std::to_string(sb) + " bytes)");
  if (ctype == 1) {
    if (!r.get(ste + 24, 8, btree) || !r.get(ste + 32, 8, heap)) return -6;
  } else {
    uint64_t ohoff;
    Dataset dummy;
    bool has_attr;
    if (!r.abs(root_oh, ohoff) || read_object(r, ohoff, dummy, btree, heap, has_attr) != OBJ_GROUP) return -6;
  }
  std::set<uint64_t> seen{root_oh};
  return walk_group(r, btree, heap, "", out, seen, 0) ? 0 : -6;
}

}  // namespace mlfh5
