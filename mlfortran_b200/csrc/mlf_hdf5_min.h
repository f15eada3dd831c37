// Minimal HDF5 writer/reader for the reference's checkpoint layout (src/mlf_hdf5.f90:671-737): root-level,
// contiguous, native little-endian int64 / float64 datasets without attributes.  No libhdf5 exists in this image,
// so the file format (superblock v0, v1 object headers, symbol-table root group) is produced directly.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace mlfh5 {

enum DType { I64 = 0, F64 = 1 };

struct Dataset {
  std::string name;
  DType type = F64;
  std::vector<uint64_t> dims;  // as stored in the file (C order == reversed Fortran order)
  std::vector<uint8_t> data;   // raw little-endian elements, 8 bytes each
  uint64_t count() const {
    uint64_t n = 1;
    for (uint64_t v : dims) n *= v;
    return n;
  }
};

// Writes a complete file holding `ds` at the root group.  Returns 0 or a negative MLF error code.
int write_file(const std::string& path, const std::vector<Dataset>& ds);
// Reads every root-level contiguous int64/float64 dataset.  Returns 0 or a negative MLF error code.
int read_file(const std::string& path, std::vector<Dataset>& out);

}  // namespace mlfh5
