#include "tmap.h"

#include <mutex>

namespace ppasr {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn(std::string* err) {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  static std::string init_err;
  std::call_once(once, [&]() {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres);
    if (e != cudaSuccess || p == nullptr || qres != cudaDriverEntryPointSuccess) {
      init_err = std::string("cudaGetDriverEntryPoint(cuTensorMapEncodeTiled) failed: ") + cudaGetErrorString(e);
      (void)cudaGetLastError();
    } else {
      fn = reinterpret_cast<EncodeTiledFn>(p);
    }
  });
  if (!fn && err) *err = init_err;
  return fn;
}

bool make_tmap_bf16_sw128(CUtensorMap* out, const void* base, int rank, const uint64_t* dims,
                          const uint64_t* strides_bytes, const uint32_t* box, std::string* err) {
  EncodeTiledFn fn = get_encode_fn(err);
  if (!fn) return false;
  cuuint64_t gdims[5];
  cuuint64_t gstrides[4];
  cuuint32_t gbox[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdims[i] = dims[i];
    gbox[i] = box[i];
    estr[i] = 1;
  }
  for (int i = 0; i + 1 < rank; ++i) gstrides[i] = strides_bytes[i];
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gdims, gstrides,
                  gbox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    if (err) {
      *err = "cuTensorMapEncodeTiled failed with CUresult " + std::to_string((int)r) + " (rank " +
             std::to_string(rank) + ", dims";
      for (int i = 0; i < rank; ++i) *err += " " + std::to_string(dims[i]);
      *err += ", box";
      for (int i = 0; i < rank; ++i) *err += " " + std::to_string(box[i]);
      *err += ")";
    }
    return false;
  }
  return true;
}

}  // namespace ppasr
