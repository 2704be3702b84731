// Host-side TMA tensor-map encoding via the driver entry point (resolved at run time through
// the CUDA runtime, so the library has no link-time dependency on libcuda).
#include "tc.cuh"
#include <mutex>

namespace tc {

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);

static EncodeTiledFn g_encode = nullptr;
static std::once_flag g_once;

static void resolve() {
  void* fn = nullptr;
  cudaDriverEntryPointQueryResult q;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q) == cudaSuccess &&
      q == cudaDriverEntryPointSuccess)
    g_encode = reinterpret_cast<EncodeTiledFn>(fn);
}

int encode_tmap(CUtensorMap* out, const void* ptr, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box) {
  std::call_once(g_once, resolve);
  if (!g_encode) {
    mp_set_error("cuTensorMapEncodeTiled is not available from this driver");
    return MP_ERR_CUDA;
  }
  cuuint64_t gdim[5], gstr[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    if (i > 0) gstr[i - 1] = strides_bytes[i];
  }
  CUresult r = g_encode(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(ptr),
                        gdim, gstr, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                        CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    mp_set_error("cuTensorMapEncodeTiled failed (CUresult %d): rank %d dims [%llu %llu %llu %llu %llu] "
                 "box [%u %u %u %u %u]",
                 (int)r, rank, (unsigned long long)gdim[0], (unsigned long long)(rank > 1 ? gdim[1] : 0),
                 (unsigned long long)(rank > 2 ? gdim[2] : 0), (unsigned long long)(rank > 3 ? gdim[3] : 0),
                 (unsigned long long)(rank > 4 ? gdim[4] : 0), bx[0], rank > 1 ? bx[1] : 0,
                 rank > 2 ? bx[2] : 0, rank > 3 ? bx[3] : 0, rank > 4 ? bx[4] : 0);
    return MP_ERR_CUDA;
  }
  return MP_OK;
}

}  // namespace tc
