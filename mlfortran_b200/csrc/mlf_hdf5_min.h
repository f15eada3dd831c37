// Minimal HDF5 writer/reader for the reference's checkpoint layout (src/mlf_hdf5.f90:671-737): contiguous, native
// little-endian int64 / float64 datasets without attributes, at the root or inside (nested) groups -- the reference
// pushes a sub-object's state into a group named after it (Hdf5PushSubObject, src/mlf_hdf5.f90:657-669; getSubGroup,
// :133-151).  No libhdf5 exists in this image, so the file format (superblock v0, v1 object headers, symbol-table groups)
// is produced directly; the reader also accepts what stock libhdf5 writes in that ("earliest") format: a user block in
// front of the superblock, header continuation blocks, attributes (ignored) and version-1/2 layout messages.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace mlfh5 {

enum DType { I64 = 0, F64 = 1 };

struct Dataset {
  std::string name;            // path below the root, components separated by '/' ("Mu", "kmeans/Mu")
  DType type = F64;
  std::vector<uint64_t> dims;  // as stored in the file (C order == reversed Fortran order)
  std::vector<uint8_t> data;   // raw little-endian elements, 8 bytes each
  uint64_t count() const {
    uint64_t n = 1;
    for (uint64_t v : dims) n *= v;
    return n;
  }
};

struct Contents {
  std::vector<Dataset> ds;
  std::vector<std::string> groups;   // every group path below the root (also the empty ones)
  // objects / features of the file that this reader cannot represent and a rewrite would therefore drop: datasets of
  // other types or layouts (chunked, compact, compressed), attributes, links it cannot follow, a user block
  std::vector<std::string> lossy;
};

// Writes a complete file (to a temporary beside `path`, then rename(2): a crash never leaves a half-written checkpoint).
// Returns 0 or a negative MLF error code.
int write_file(const std::string& path, const Contents& c);
// Reads every contiguous int64/float64 dataset of every group.  Returns 0 or a negative MLF error code; malformed or
// truncated files are rejected (every field is bounds-checked), never read out of bounds.
int read_file(const std::string& path, Contents& out);

}  // namespace mlfh5
