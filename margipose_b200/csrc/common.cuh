// Shared helpers for the margipose_b200 CUDA kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>

#define MP_OK 0
#define MP_ERR_ARG (-1)       // invalid argument (shape / alignment / null pointer)
#define MP_ERR_CUDA (-2)      // a CUDA runtime call failed; see mp_last_error()
#define MP_ERR_UNSUPPORTED (-3)

void mp_set_error(const char* fmt, ...);

#define MP_CHECK_ARG(cond, ...)            \
  do {                                     \
    if (!(cond)) {                         \
      mp_set_error(__VA_ARGS__);           \
      return MP_ERR_ARG;                   \
    }                                      \
  } while (0)

#define MP_CHECK_LAUNCH(name)                                              \
  do {                                                                     \
    cudaError_t e__ = cudaGetLastError();                                  \
    if (e__ != cudaSuccess) {                                              \
      mp_set_error("%s: launch failed: %s", name, cudaGetErrorString(e__)); \
      return MP_ERR_CUDA;                                                  \
    }                                                                      \
  } while (0)

#define MP_CUDA(call)                                                          \
  do {                                                                         \
    cudaError_t e__ = (call);                                                  \
    if (e__ != cudaSuccess) {                                                  \
      mp_set_error("%s failed: %s", #call, cudaGetErrorString(e__));           \
      return MP_ERR_CUDA;                                                      \
    }                                                                          \
  } while (0)

// Programmatic dependent launch (tunable "pdl", default on): a kernel launched through mp_launch may
// be scheduled while its stream predecessor is still running -- its CTAs become resident and run their
// prologue early -- and must call pdl_wait() before it touches memory the predecessor writes or reads;
// pdl_trigger() lets the NEXT kernel do the same.  Both are no-ops in a normally launched kernel.
int mp_pdl_enabled();
// Tunable "deterministic": kernels that combine partial sums with floating-point atomics switch to a fixed order
// (see include/margipose_b200.h); the launchers read it at launch time.
int mp_deterministic();
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
#ifdef __CUDACC__
template <typename... KArgs, typename... Args>
static inline cudaError_t mp_launch(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                    Args... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = mp_pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

static inline bool mp_aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

__device__ __forceinline__ float bf16_round(float x) { return __bfloat162float(__float2bfloat16_rn(x)); }

__device__ __forceinline__ uint32_t pack_bf16x2(float lo, float hi) {
  __nv_bfloat162 v = __floats2bfloat162_rn(lo, hi);
  return *reinterpret_cast<uint32_t*>(&v);
}
__device__ __forceinline__ float2 unpack_bf16x2(uint32_t u) {
  __nv_bfloat162 v = *reinterpret_cast<__nv_bfloat162*>(&u);
  return __bfloat1622float2(v);
}

// 8 consecutive channels of an activation tensor as fp32.  lo_delta == 0: plain bf16.  lo_delta > 0 (split / "bf16x3"
// mode): the value is the sum of two bf16 tensors, hi at p and lo at p + lo_delta (hi = bf16(v), lo = bf16(v - hi)).
__device__ __forceinline__ void mp_ld8(const __nv_bfloat16* p, long long lo_delta, float (&v)[8]) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  float2 f;
  f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
  f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
  f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
  f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
  if (lo_delta) {
    const uint4 l = __ldg(reinterpret_cast<const uint4*>(p + lo_delta));
    f = unpack_bf16x2(l.x); v[0] += f.x; v[1] += f.y;
    f = unpack_bf16x2(l.y); v[2] += f.x; v[3] += f.y;
    f = unpack_bf16x2(l.z); v[4] += f.x; v[5] += f.y;
    f = unpack_bf16x2(l.w); v[6] += f.x; v[7] += f.y;
  }
}
__device__ __forceinline__ void mp_st8(__nv_bfloat16* p, long long lo_delta, const float (&v)[8]) {
  const uint4 hi = make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                              pack_bf16x2(v[6], v[7]));
  *reinterpret_cast<uint4*>(p) = hi;
  if (lo_delta) {
    const float2 a = unpack_bf16x2(hi.x), b = unpack_bf16x2(hi.y), c = unpack_bf16x2(hi.z), d = unpack_bf16x2(hi.w);
    *reinterpret_cast<uint4*>(p + lo_delta) =
        make_uint4(pack_bf16x2(v[0] - a.x, v[1] - a.y), pack_bf16x2(v[2] - b.x, v[3] - b.y),
                   pack_bf16x2(v[4] - c.x, v[5] - c.y), pack_bf16x2(v[6] - d.x, v[7] - d.y));
  }
}
