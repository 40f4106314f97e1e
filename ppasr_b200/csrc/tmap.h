// Host-side CUtensorMap construction. cuTensorMapEncodeTiled is resolved at run time through
// cudaGetDriverEntryPoint so the shared library has no link-time dependency on libcuda.so
// (it must load, and export its symbols, on a machine without a GPU driver).
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>
#include <string>

namespace ppasr {

// Encodes a bf16 tensor map with the 128B swizzle. dims[0] is the innermost (contiguous) dimension.
// strides_bytes[i] is the byte stride of dims[i+1]. box[0] * 2 bytes must equal 128.
// Returns false and fills `err` on failure.
bool make_tmap_bf16_sw128(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, std::string* err);

inline bool make_tmap_2d(CUtensorMap* out, const void* base, uint64_t inner, uint64_t outer, uint64_t pitch_bytes,
                         uint32_t box_outer, std::string* err) {
  uint64_t dims[2] = {inner, outer};
  uint64_t strides[1] = {pitch_bytes};
  uint32_t box[2] = {64, box_outer};
  return make_tmap_bf16_sw128(out, base, 2, dims, strides, box, err);
}

inline bool make_tmap_3d(CUtensorMap* out, const void* base, uint64_t d0, uint64_t d1, uint64_t d2, uint64_t s1_bytes,
                         uint64_t s2_bytes, uint32_t box1, std::string* err) {
  uint64_t dims[3] = {d0, d1, d2};
  uint64_t strides[2] = {s1_bytes, s2_bytes};
  uint32_t box[3] = {64, box1, 1};
  return make_tmap_bf16_sw128(out, base, 3, dims, strides, box, err);
}

}  // namespace ppasr
