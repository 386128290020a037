// Checkpoint container: an uncompressed ``.npz`` (a stored ZIP of ``.npy`` members), the format ``numpy.savez`` writes
// and ``numpy.load`` reads.  Host code only; used by drs_save / drs_load (include/drs.h) and by the drs_npz_* entry
// points.  Replaces ``tf.train.Saver.save / restore`` of the reference (isprs:1693-1717, 1797-1802) with a file that can be
// inspected, converted and produced with NumPy alone.
#pragma once
#include <cstdint>
#include <string>
#include <vector>

namespace drs_npz {

struct Array {
  std::string name;            // member name without ".npy"
  std::vector<int64_t> shape;  // C order
  std::vector<float> data;     // always float32 in memory
  int64_t count() const {
    int64_t n = 1;
    for (int64_t d : shape) n *= d;
    return n;
  }
};

// Both return an empty string on success, else the message for drs_last_error.
// write(): members in the given order, float32 little-endian, written to "<path>.tmp.<pid>" and renamed over <path>.
std::string write(const std::string& path, const std::vector<Array>& arrays);
// read(): every member of a stored (uncompressed) archive; <f4, <f8, <i4, <i8, <u1 payloads are converted to float32.
// Compressed members (numpy.savez_compressed), Fortran order and big-endian payloads are refused.
std::string read(const std::string& path, std::vector<Array>& arrays);

uint32_t crc32(const void* data, size_t n, uint32_t crc = 0);

}  // namespace drs_npz
