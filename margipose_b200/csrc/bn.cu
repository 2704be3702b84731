// BatchNorm + ReLU + residual, forward and backward, on bf16 NHWC activations (sm_100a).
//
// Replaces nn.BatchNorm2d / nn.ReLU / the residual `+` of ResidualBlock
// (/root/reference/src/margipose/models/margipose_model.py:31-40) and of the torchvision ResNet
// blocks (:130-135), and their autograd.  HBM/L2-bound elementwise work: every tensor is touched
// once per pass with 128-bit loads/stores (8 bf16 channels per thread, U pixels in flight per
// thread); the batch statistics arrive pre-reduced from the conv epilogue (igemm.cu), so the forward
// is a single pass.  The backward is two passes (per-channel reductions, then the gradient), the
// minimum for training-mode BatchNorm.  Algorithmic bytes per pixel-channel: forward 2 per input
// tensor + 2 out; backward reduce 2 per tensor read; apply 2 per tensor read + 2 per gradient.
#include "tc.cuh"
#include "../../include/margipose_b200.h"

// tunable "bn_tma": bit mask of the kernels that use their TMA-staged variant where it applies (plain bf16 NHWC):
// 1 = forward, 2 = backward reduce, 4 = backward apply.  Default 1: in the training step the backward kernels run
// beside the weight-gradient kernels of the auxiliary stream (200 KB of shared memory per CTA), and a 100 KB tile ring
// cannot share an SM with those -- measured: backward program 7.80 ms register-staged vs 8.13 ms TMA-staged, although
// each TMA kernel alone is 2-5 % faster (tools/step_time.py, tools/bn_time.py).
long long g_bn_tma = 1;

namespace {

constexpr int MAXT = 256;

struct BnGroup {   // problems of a grouped launch (blockIdx.y): same shapes and flags, different tensors
  mp_bn_args a[MP_MAX_GROUP];
};
constexpr int U = 2;    // pixels per thread per iteration (forward, backward reduce); the NEXT iteration's
constexpr int UA = 1;   // loads are issued before the current one is consumed (software pipelining)

__device__ __forceinline__ void unpack8(const uint4& u, float (&v)[8]) {
  float2 f;
  f = unpack_bf16x2(u.x); v[0] = f.x; v[1] = f.y;
  f = unpack_bf16x2(u.y); v[2] = f.x; v[3] = f.y;
  f = unpack_bf16x2(u.z); v[4] = f.x; v[5] = f.y;
  f = unpack_bf16x2(u.w); v[6] = f.x; v[7] = f.y;
}
__device__ __forceinline__ uint4 pack8(const float (&v)[8]) {
  return make_uint4(pack_bf16x2(v[0], v[1]), pack_bf16x2(v[2], v[3]), pack_bf16x2(v[4], v[5]),
                    pack_bf16x2(v[6], v[7]));
}
__device__ __forceinline__ uint4 ldg16(const void* base, long long elem_off) {
  return __ldg(reinterpret_cast<const uint4*>(reinterpret_cast<const __nv_bfloat16*>(base) + elem_off));
}

// An 8-channel group of an activation tensor: one 128-bit load (plain bf16) or two (split mode: value = hi + lo,
// lo stored lo_delta elements after hi).
template <bool SPLIT>
struct Act8 {
  uint4 hi;
};
template <>
struct Act8<true> {
  uint4 hi, lo;
};
template <bool SPLIT>
__device__ __forceinline__ Act8<SPLIT> ld_act(const void* base, long long off, long long lo_delta) {
  Act8<SPLIT> a;
  a.hi = ldg16(base, off);
  if constexpr (SPLIT) a.lo = ldg16(base, off + lo_delta);
  return a;
}
template <bool SPLIT>
__device__ __forceinline__ void act_unpack(const Act8<SPLIT>& a, float (&v)[8]) {
  unpack8(a.hi, v);
  if constexpr (SPLIT) {
    float t[8];
    unpack8(a.lo, t);
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] += t[j];
  }
}
template <bool SPLIT>
__device__ __forceinline__ void st_act(void* base, long long off, long long lo_delta, const float (&v)[8]) {
  const uint4 hi = pack8(v);
  *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + off) = hi;
  if constexpr (SPLIT) {
    float h[8], r[8];
    unpack8(hi, h);
#pragma unroll
    for (int j = 0; j < 8; ++j) r[j] = v[j] - h[j];
    *reinterpret_cast<uint4*>(reinterpret_cast<__nv_bfloat16*>(base) + off + lo_delta) = pack8(r);
  }
}

constexpr int MAXR = 8;   // replicas summed with all loads in flight (more fall back to a loop)

// Sum over replicas of 8 consecutive channels of a per-channel statistic: 2 x 128-bit loads per
// replica, all issued before the first add (one L2 round trip instead of a dependent chain).
__device__ __forceinline__ void rsum8(const float* p, int c0, int replicas, long long stride, float (&out)[8]) {
  float4 lo[MAXR], hi[MAXR];
#pragma unroll
  for (int r = 0; r < MAXR; ++r) {
    if (r < replicas) {
      const float4* q = reinterpret_cast<const float4*>(p + (long long)r * stride + c0);
      lo[r] = __ldg(q);
      hi[r] = __ldg(q + 1);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) out[j] = 0.f;
#pragma unroll
  for (int r = 0; r < MAXR; ++r) {
    if (r < replicas) {
      out[0] += lo[r].x; out[1] += lo[r].y; out[2] += lo[r].z; out[3] += lo[r].w;
      out[4] += hi[r].x; out[5] += hi[r].y; out[6] += hi[r].z; out[7] += hi[r].w;
    }
  }
  for (int r = MAXR; r < replicas; ++r)
    for (int j = 0; j < 8; ++j) out[j] += p[(long long)r * stride + c0 + j];
}

// mean / invstd / biased variance of 8 channels: forward from the conv-epilogue sums (training) or
// the running buffers (eval); backward from what the forward saved (bitwise the same numbers).
__device__ __forceinline__ void channel_stats8(const mp_bn_args& A, const mp_bn_branch& br, int c0, bool from_saved,
                                               float (&mean)[8], float (&invstd)[8], float (&var)[8]) {
  if (!from_saved && A.training) {
    float s1[8], s2[8];
    rsum8(br.sum, c0, A.stat_replicas, A.stat_stride, s1);
    rsum8(br.sq, c0, A.stat_replicas, A.stat_stride, s2);
    const float inv_m = 1.0f / (float)A.M;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mean[i] = s1[i] * inv_m;
      var[i] = fmaxf(s2[i] * inv_m - mean[i] * mean[i], 0.f);
      invstd[i] = rsqrtf(var[i] + A.eps);
    }
    return;
  }
  // (every load below is issued before the first use: a guarded load followed by its use, eight times over, is a chain
  // of eight dependent L2 / DRAM round trips in front of the block's streaming loop -- tools/bn_scale.py from_sums)
  if (from_saved) {   // Cp-sized buffers: two 128-bit loads each
    const float4 m0 = __ldg(reinterpret_cast<const float4*>(br.save_mean + c0)), m1 = __ldg(reinterpret_cast<const float4*>(br.save_mean + c0) + 1);
    const float4 i0 = __ldg(reinterpret_cast<const float4*>(br.save_invstd + c0)), i1 = __ldg(reinterpret_cast<const float4*>(br.save_invstd + c0) + 1);
    mean[0] = m0.x; mean[1] = m0.y; mean[2] = m0.z; mean[3] = m0.w; mean[4] = m1.x; mean[5] = m1.y; mean[6] = m1.z; mean[7] = m1.w;
    invstd[0] = i0.x; invstd[1] = i0.y; invstd[2] = i0.z; invstd[3] = i0.w;
    invstd[4] = i1.x; invstd[5] = i1.y; invstd[6] = i1.z; invstd[7] = i1.w;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      var[i] = 0.f;
      if (c0 + i >= A.C) mean[i] = invstd[i] = 0.f;
    }
    return;
  }
  float rm[8], rv[8], cb[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {   // C-sized module buffers: predicated scalar loads, all in flight together
    const bool ok = c0 + i < A.C;
    rm[i] = ok ? br.running_mean[c0 + i] : 0.f;
    rv[i] = ok ? br.running_var[c0 + i] : 1.f;
    cb[i] = (ok && br.conv_bias) ? br.conv_bias[c0 + i] : 0.f;
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool ok = c0 + i < A.C;
    mean[i] = ok ? rm[i] - cb[i] : 0.f;
    invstd[i] = ok ? rsqrtf(rv[i] + A.eps) : 0.f;
    var[i] = 0.f;
  }
}

__device__ __forceinline__ void load8f(const float* p, int c0, float (&v)[8]) {
  const float4 lo = __ldg(reinterpret_cast<const float4*>(p + c0));
  const float4 hi = __ldg(reinterpret_cast<const float4*>(p + c0) + 1);
  v[0] = lo.x; v[1] = lo.y; v[2] = lo.z; v[3] = lo.w; v[4] = hi.x; v[5] = hi.y; v[6] = hi.z; v[7] = hi.w;
}

__device__ __forceinline__ void affine(const mp_bn_args& A, const mp_bn_branch& br, int c0, bool from_saved,
                                       float (&scale)[8], float (&shift)[8]) {
  if (br.scale && A.training) {   // finalised by the producing conv's last CTA (igemm.cu)
    load8f(br.scale, c0, scale);
    load8f(br.shift, c0, shift);
    return;
  }
  float gv[8], bv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {   // C-sized parameters: predicated loads, issued before anything waits on the statistics
    const bool ok = c0 + i < A.C;
    gv[i] = ok ? br.gamma[c0 + i] : 0.f;
    bv[i] = ok ? br.beta[c0 + i] : 0.f;
  }
  float mean[8], invstd[8], var[8];
  channel_stats8(A, br, c0, from_saved, mean, invstd, var);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    scale[i] = gv[i] * invstd[i];                 // zero beyond C
    shift[i] = fmaf(-mean[i], scale[i], bv[i]);
  }
}

// Saved statistics + running buffers of one BatchNorm (one thread row of one block per launch).  A real function call
// with the branch's pointers passed BY VALUE: inlined into the kernels it costs them hundreds of bytes of spills, and
// through a reference to the __grid_constant__ argument struct every field access is a generic load of parameter space.
// Every load is issued before the first store (interleaved read-modify-writes form a chain of dependent round trips).
struct BookArgs {
  const float* sum; const float* sq; const float* conv_bias;
  float* running_mean; float* running_var; float* save_mean; float* save_invstd;
  long long M, stat_stride;
  int C, stat_replicas, training;
  float momentum, eps;
};
__device__ __noinline__ void bookkeeping_call(const BookArgs B, const int c0) {
  float mean[8], invstd[8], var[8], rm[8], rv[8], bias[8];
  const bool run = B.training && B.running_mean;
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const bool ok = c0 + i < B.C && (run || !B.training);
    rm[i] = ok ? B.running_mean[c0 + i] : 0.f;
    rv[i] = ok ? B.running_var[c0 + i] : 1.f;
    bias[i] = (ok && B.conv_bias) ? B.conv_bias[c0 + i] : 0.f;
  }
  if (B.training) {
    float s1[8], s2[8];
    rsum8(B.sum, c0, B.stat_replicas, B.stat_stride, s1);
    rsum8(B.sq, c0, B.stat_replicas, B.stat_stride, s2);
    const float inv_m = 1.0f / (float)B.M;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mean[i] = s1[i] * inv_m;
      var[i] = fmaxf(s2[i] * inv_m - mean[i] * mean[i], 0.f);
      invstd[i] = rsqrtf(var[i] + B.eps);
    }
  } else {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      mean[i] = rm[i] - bias[i];
      invstd[i] = rsqrtf(rv[i] + B.eps);
      var[i] = 0.f;
    }
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + i;
    if (c >= B.C) continue;
    if (B.save_mean) {
      B.save_mean[c] = mean[i];
      B.save_invstd[c] = invstd[i];
    }
    if (run) {
      const float unbiased = B.M > 1 ? var[i] * ((float)B.M / (float)(B.M - 1)) : var[i];
      B.running_mean[c] = (1.f - B.momentum) * rm[i] + B.momentum * (mean[i] + bias[i]);
      B.running_var[c] = (1.f - B.momentum) * rv[i] + B.momentum * unbiased;
    }
  }
}
// (A and br are accessed statically by the caller: constant-bank reads)
#define MP_BN_BOOKKEEPING(A, br, c0)                                                                              \
  do {                                                                                                            \
    if (!((br).scale && (A).training)) {   /* otherwise the producing conv already did it */                      \
      BookArgs B__;                                                                                               \
      B__.sum = (br).sum; B__.sq = (br).sq; B__.conv_bias = (br).conv_bias;                                       \
      B__.running_mean = (br).running_mean; B__.running_var = (br).running_var;                                   \
      B__.save_mean = (br).save_mean; B__.save_invstd = (br).save_invstd;                                         \
      B__.M = (A).M; B__.stat_stride = (A).stat_stride; B__.C = (A).C; B__.stat_replicas = (A).stat_replicas;     \
      B__.training = (A).training; B__.momentum = (A).momentum; B__.eps = (A).eps;                                \
      bookkeeping_call(B__, c0);                                                                                  \
    }                                                                                                             \
  } while (0)

// ------------------------------------------------------------------------------------- forward
template <bool SPLIT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) bn_fwd_kernel(const __grid_constant__ BnGroup GRP) {
  pdl_trigger();
  pdl_wait();
  const mp_bn_args& A = GRP.a[blockIdx.y];
  const int c0 = threadIdx.x * 8;
  const bool has_b = A.b.y != nullptr;
  float sa[8], ha[8], sb[8], hb[8];
  affine(A, A.a, c0, false, sa, ha);
  if (has_b) affine(A, A.b, c0, false, sb, hb);

  if (blockIdx.x == 0 && threadIdx.y == 0) {   // bookkeeping: saved statistics + running buffers
    MP_BN_BOOKKEEPING(A, A.a, c0);
    if (has_b) MP_BN_BOOKKEEPING(A, A.b, c0);
  }

  const long long ppb = (long long)blockDim.y * U;
  const long long stride = (long long)gridDim.x * ppb;
  auto load = [&](long long p0, Act8<SPLIT> (&la)[U], Act8<SPLIT> (&lb)[U]) {
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const long long pix = p0 + (long long)i * blockDim.y;
      if (pix < A.M) {
        const long long off = pix * A.Cp + c0;
        la[i] = ld_act<SPLIT>(A.a.y, off, A.lo_delta);
        if (has_b) lb[i] = ld_act<SPLIT>(A.b.y, off, A.lo_delta);
        else if (A.res) lb[i] = ld_act<SPLIT>(A.res, off, A.lo_delta);
      }
    }
  };
  Act8<SPLIT> la[U], lb[U], na[U], nb[U];
  long long base = (long long)blockIdx.x * ppb;
  if (base < A.M) load(base + threadIdx.y, la, lb);
  for (; base < A.M; base += stride) {
    if (base + stride < A.M) load(base + stride + threadIdx.y, na, nb);
    const long long p0 = base + threadIdx.y;
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const long long pix = p0 + (long long)i * blockDim.y;
      if (pix >= A.M) break;
      const long long off = pix * A.Cp + c0;
      float z[8], t[8];
      act_unpack<SPLIT>(la[i], z);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        z[j] = fmaf(z[j], sa[j], ha[j]);
        if (A.relu_a) z[j] = fmaxf(z[j], 0.f);
      }
      if (has_b) {
        act_unpack<SPLIT>(lb[i], t);
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] += fmaf(t[j], sb[j], hb[j]);
      } else if (A.res) {
        act_unpack<SPLIT>(lb[i], t);
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] += t[j];
      }
      if (A.relu_out) {
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] = fmaxf(z[j], 0.f);
      }
      if (c0 + 8 > A.C) {   // only the thread that owns the partially filled channel group masks its padding
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j >= A.C) z[j] = 0.f;
      }
      if (A.out) st_act<SPLIT>(A.out, off, A.lo_delta, z);
      if (A.out_nchw) {
        const long long n = pix / A.HW, hw = pix - n * A.HW;
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j < A.C) A.out_nchw[(n * A.C + c0 + j) * A.HW + hw] = z[j];
      }
    }
#pragma unroll
    for (int i = 0; i < U; ++i) { la[i] = na[i]; lb[i] = nb[i]; }
  }
}

// ------------------------------------------------------------------------------------ backward
// Gradient w.r.t. the two BN outputs at one pixel: g masked by the post-ReLU (out > 0) and, for
// branch a, by the pre-ReLU (bn_a(y_a) > 0, recomputed with the forward's exact coefficients).
// NCHW = the incoming gradient is the fp32 (N, C, HW) logits gradient (last block of a column);
// its 8 channel values then travel in the two 128-bit slots g / g2 instead of one bf16x8 load.
template <bool SPLIT>
struct Loads {
  Act8<SPLIT> g, ya, yb;   // NCHW: g.hi carries the first four fp32 gradient values, g2 the other four
  uint4 o;                 // forward output, read for the sign of the post-ReLU only (hi decides it)
};
template <bool NCHW, bool SPLIT>
struct LoadsX : Loads<SPLIT> {};
template <bool SPLIT>
struct LoadsX<true, SPLIT> : Loads<SPLIT> {
  uint4 g2;
};

template <bool NCHW, bool SPLIT>
__device__ __forceinline__ void load_pixel(const mp_bn_args& A, bool has_b, long long pix, int c0,
                                           LoadsX<NCHW, SPLIT>& L) {
  const long long off = pix * A.Cp + c0;
  if constexpr (!NCHW) {
    L.g = ld_act<SPLIT>(A.dout, off, A.lo_delta);
  } else {
    const long long n = pix / A.HW, hw = pix - n * A.HW;
    float gn[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) gn[j] = (c0 + j < A.C) ? __ldg(A.dout_nchw + (n * A.C + c0 + j) * A.HW + hw) : 0.f;
    L.g.hi = make_uint4(__float_as_uint(gn[0]), __float_as_uint(gn[1]), __float_as_uint(gn[2]), __float_as_uint(gn[3]));
    L.g2 = make_uint4(__float_as_uint(gn[4]), __float_as_uint(gn[5]), __float_as_uint(gn[6]), __float_as_uint(gn[7]));
  }
  if (A.relu_out) L.o = ldg16(A.out, off);
  L.ya = ld_act<SPLIT>(A.a.y, off, A.lo_delta);
  if (has_b) L.yb = ld_act<SPLIT>(A.b.y, off, A.lo_delta);
}

template <bool NCHW, bool SPLIT>
__device__ __forceinline__ void grads_at(const mp_bn_args& A, const LoadsX<NCHW, SPLIT>& L, const float (&sa)[8],
                                         const float (&ha)[8], int c0, float (&dza)[8], float (&dzb)[8],
                                         float (&ya)[8]) {
  float g[8];
  if constexpr (!NCHW) {
    act_unpack<SPLIT>(L.g, g);
  } else {
    g[0] = __uint_as_float(L.g.hi.x); g[1] = __uint_as_float(L.g.hi.y); g[2] = __uint_as_float(L.g.hi.z);
    g[3] = __uint_as_float(L.g.hi.w); g[4] = __uint_as_float(L.g2.x); g[5] = __uint_as_float(L.g2.y);
    g[6] = __uint_as_float(L.g2.z); g[7] = __uint_as_float(L.g2.w);
  }
  if (A.relu_out) {
    float o[8];
    unpack8(L.o, o);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (!(o[j] > 0.f)) g[j] = 0.f;
  }
  act_unpack<SPLIT>(L.ya, ya);
  if (c0 + 8 > A.C) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c0 + j >= A.C) g[j] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    dzb[j] = g[j];
    dza[j] = (A.relu_a && !(fmaf(ya[j], sa[j], ha[j]) > 0.f)) ? 0.f : g[j];
  }
}

// sums layout per replica: [0] sum dz_a, [1] sum dz_a * y_a, [2] sum dz_b, [3] sum dz_b * y_b
template <bool NCHW, bool SPLIT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) bn_bwd_reduce_kernel(const __grid_constant__ BnGroup GRP) {
  pdl_trigger();
  pdl_wait();
  const mp_bn_args& A = GRP.a[blockIdx.y];
  __shared__ float red[32 * MAXT];
  const int cg = threadIdx.x, c0 = cg * 8;
  const bool has_b = A.b.y != nullptr;
  float sa[8], ha[8];
  affine(A, A.a, c0, true, sa, ha);
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;

  const long long ppb = (long long)blockDim.y * U;
  const long long stride = (long long)gridDim.x * ppb;
  auto load = [&](long long p0, LoadsX<NCHW, SPLIT> (&L)[U]) {
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const long long pix = p0 + (long long)i * blockDim.y;
      if (pix < A.M) load_pixel<NCHW, SPLIT>(A, has_b, pix, c0, L[i]);
    }
  };
  LoadsX<NCHW, SPLIT> L[U], Ln[U];
  long long base = (long long)blockIdx.x * ppb;
  if (base < A.M) load(base + threadIdx.y, L);
  for (; base < A.M; base += stride) {
    if (base + stride < A.M) load(base + stride + threadIdx.y, Ln);
    const long long p0 = base + threadIdx.y;
#pragma unroll
    for (int i = 0; i < U; ++i) {
      const long long pix = p0 + (long long)i * blockDim.y;
      if (pix >= A.M) break;
      float dza[8], dzb[8], ya[8], yb[8];
      grads_at<NCHW, SPLIT>(A, L[i], sa, ha, c0, dza, dzb, ya);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j] += dza[j];
        acc[8 + j] = fmaf(dza[j], ya[j], acc[8 + j]);
      }
      if (has_b) {
        act_unpack<SPLIT>(L[i].yb, yb);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[16 + j] += dzb[j];
          acc[24 + j] = fmaf(dzb[j], yb[j], acc[24 + j]);
        }
      }
    }
#pragma unroll
    for (int i = 0; i < U; ++i) L[i] = Ln[i];
  }
  // block reduction over the pixel rows, all threads take part; then one atomic per (sum, channel)
  const int G = blockDim.x, PY = blockDim.y;
  const int tid = threadIdx.y * G + cg;
  const int nsum = has_b ? 32 : 16;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (j < nsum) red[j * MAXT + tid] = acc[j];
  __syncthreads();
  float* dst = A.sums + (long long)(blockIdx.x % A.stat_replicas) * 4 * A.Cp;
  for (int j = threadIdx.y; j < nsum; j += PY) {
    float s = 0.f;
    for (int y = 0; y < PY; ++y) s += red[j * MAXT + y * G + cg];
    atomicAdd(dst + (j >> 3) * A.Cp + c0 + (j & 7), s);
  }
  if (A.bwd_counter == nullptr) return;
  // Last block to arrive turns the sums into the per-channel coefficients of the apply pass
  // (dy = coef0*dz + coef1*y + coef2) and accumulates the affine-parameter gradients.
  __shared__ unsigned s_ticket;
  __threadfence();
  __syncthreads();
  if (tid == 0) s_ticket = atomicAdd(A.bwd_counter, 1u);
  __syncthreads();
  if (s_ticket != gridDim.x - 1) return;
  __threadfence();
  const float inv_m = 1.0f / (float)A.M;
  for (int c = tid; c < A.Cp; c += G * PY) {
    for (int which = 0; which < (has_b ? 2 : 1); ++which) {
      const mp_bn_branch& br = which ? A.b : A.a;
      float k0 = 0.f, k1 = 0.f, k2 = 0.f;
      if (c < A.C) {
        const float mean = br.save_mean[c], invstd = br.save_invstd[c];
        const float s1 = __ldcg(A.sums + (2 * which) * A.Cp + c), s2 = __ldcg(A.sums + (2 * which + 1) * A.Cp + c);
        const float dgamma = invstd * (s2 - mean * s1);
        k0 = br.gamma[c] * invstd;
        k1 = -k0 * dgamma * inv_m * invstd;
        k2 = -k0 * s1 * inv_m - k1 * mean;
        if (br.dbeta) br.dbeta[c] += s1;
        if (br.dgamma) br.dgamma[c] += dgamma;
      }
      br.coef[c] = k0;
      br.coef[A.Cp + c] = k1;
      br.coef[2 * A.Cp + c] = k2;
    }
  }
}

// dy = scale * (dz - mean(dz) - x^ * mean(dz * x^)) = scale*dz + kb*y + kc per channel
__device__ __forceinline__ void bwd_coefs(const mp_bn_args& A, const mp_bn_branch& br, int which, int c0,
                                          float (&scale)[8], float (&kb)[8], float (&kc)[8], bool write_param_grads) {
  const float inv_m = 1.0f / (float)A.M;
  float s1v[8], s2v[8], meanv[8], invv[8];
  rsum8(A.sums + (2 * which) * A.Cp, c0, A.stat_replicas, 4 * A.Cp, s1v);
  rsum8(A.sums + (2 * which + 1) * A.Cp, c0, A.stat_replicas, 4 * A.Cp, s2v);
  load8f(br.save_mean, c0, meanv);      // (Cp)-sized buffers: vector loads, all in flight together
  load8f(br.save_invstd, c0, invv);
  float gv[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) gv[i] = (c0 + i < A.C) ? __ldg(br.gamma + c0 + i) : 0.f;   // (C-sized parameter: guarded scalar loads, no stores in between)
  float dg[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + i;
    scale[i] = kb[i] = kc[i] = dg[i] = 0.f;
    if (c >= A.C) continue;
    const float mean = meanv[i], invstd = invv[i];
    const float s1 = s1v[i], s2 = s2v[i];
    const float dgamma = invstd * (s2 - mean * s1);     // sum dz * x^
    const float sc = gv[i] * invstd;
    scale[i] = sc;
    kb[i] = -sc * dgamma * inv_m * invstd;
    kc[i] = -sc * s1 * inv_m - kb[i] * mean;
    dg[i] = dgamma;
  }
  if (write_param_grads) {
    // One thread per channel and launch adds to the parameter gradients, so the sum is still deterministic; the
    // result-less atomics (RED) cost no load latency -- a read-modify-write chain here (16 dependent L2 / DRAM round
    // trips) held this thread row's share of the streaming loop back by ~10 us per launch.
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      if (c0 + i >= A.C) continue;
      if (br.dbeta) atomicAdd(br.dbeta + c0 + i, s1v[i]);
      if (br.dgamma) atomicAdd(br.dgamma + c0 + i, dg[i]);
    }
  }
}

template <bool NCHW, bool SPLIT, int MINB>
__global__ void __launch_bounds__(MAXT, MINB) bn_bwd_apply_kernel(const __grid_constant__ BnGroup GRP) {
  pdl_trigger();
  pdl_wait();
  const mp_bn_args& A = GRP.a[blockIdx.y];
  const int c0 = threadIdx.x * 8;
  const bool has_b = A.b.y != nullptr;
  const bool first = blockIdx.x == 0 && threadIdx.y == 0;
  float sa[8], ha[8], ka[8], ca[8], sb[8], kb[8], cb[8];
  affine(A, A.a, c0, true, sa, ha);
  if (A.a.coef && A.bwd_counter) {   // finalised by the last block of the reduce pass
    load8f(A.a.coef + A.Cp, c0, ka);
    load8f(A.a.coef + 2 * A.Cp, c0, ca);
    if (has_b) {
      load8f(A.b.coef, c0, sb);
      load8f(A.b.coef + A.Cp, c0, kb);
      load8f(A.b.coef + 2 * A.Cp, c0, cb);
    }
  } else {
    bwd_coefs(A, A.a, 0, c0, sa, ka, ca, first);
    if (has_b) bwd_coefs(A, A.b, 1, c0, sb, kb, cb, first);
  }

  const long long ppb = (long long)blockDim.y * UA;
  const long long stride = (long long)gridDim.x * ppb;
  auto load = [&](long long p0, LoadsX<NCHW, SPLIT> (&L)[UA]) {
#pragma unroll
    for (int i = 0; i < UA; ++i) {
      const long long pix = p0 + (long long)i * blockDim.y;
      if (pix < A.M) load_pixel<NCHW, SPLIT>(A, has_b, pix, c0, L[i]);
    }
  };
  LoadsX<NCHW, SPLIT> L[UA], Ln[UA];
  long long base = (long long)blockIdx.x * ppb;
  if (base < A.M) load(base + threadIdx.y, L);
  for (; base < A.M; base += stride) {
    if (base + stride < A.M) load(base + stride + threadIdx.y, Ln);
    const long long p0 = base + threadIdx.y;
#pragma unroll
    for (int i = 0; i < UA; ++i) {
      const long long pix = p0 + (long long)i * blockDim.y;
      if (pix >= A.M) break;
      const long long off = pix * A.Cp + c0;
      float dza[8], dzb[8], ya[8], o[8];
      grads_at<NCHW, SPLIT>(A, L[i], sa, ha, c0, dza, dzb, ya);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(sa[j], dza[j], fmaf(ka[j], ya[j], ca[j]));   // coefficients are 0 beyond C
      if (A.a.dy) st_act<SPLIT>(A.a.dy, off, A.lo_delta, o);
      if (has_b) {
        float yb[8];
        act_unpack<SPLIT>(L[i].yb, yb);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(sb[j], dzb[j], fmaf(kb[j], yb[j], cb[j]));
        if (A.b.dy) st_act<SPLIT>(A.b.dy, off, A.lo_delta, o);
      } else if (A.dres) {
        st_act<SPLIT>(A.dres, off, A.lo_delta, dzb);
      }
    }
#pragma unroll
    for (int i = 0; i < UA; ++i) L[i] = Ln[i];
  }
}

// Batch statistics of one conv output as a separate, run-to-run deterministic pass (deterministic mode: the conv
// epilogue's atomically accumulated sums are not used).  Block r of a problem sums pixels [r*M/R, (r+1)*M/R) in a
// fixed order and STORES its partial sums into replica r of sum / sq; mp_bn_fwd adds the replicas in order.
template <bool SPLIT>
__global__ void __launch_bounds__(MAXT) bn_stats_kernel(const __grid_constant__ BnGroup GRP) {
  pdl_trigger();
  pdl_wait();
  const mp_bn_args& A = GRP.a[blockIdx.y];
  __shared__ float red[16 * MAXT];
  const int cg = threadIdx.x, c0 = cg * 8;
  const int G = blockDim.x, PY = blockDim.y;
  const int R = gridDim.x;
  const long long lo = (long long)blockIdx.x * A.M / R, hi = (long long)(blockIdx.x + 1) * A.M / R;
  float acc[16];
#pragma unroll
  for (int j = 0; j < 16; ++j) acc[j] = 0.f;
  for (long long pix = lo + threadIdx.y; pix < hi; pix += PY) {
    float v[8];
    act_unpack<SPLIT>(ld_act<SPLIT>(A.a.y, pix * A.Cp + c0, A.lo_delta), v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[j] += v[j];
      acc[8 + j] = fmaf(v[j], v[j], acc[8 + j]);
    }
  }
  const int tid = threadIdx.y * G + cg;
#pragma unroll
  for (int j = 0; j < 16; ++j) red[j * MAXT + tid] = acc[j];
  __syncthreads();
  float* sum = const_cast<float*>(A.a.sum) + (long long)blockIdx.x * A.stat_stride;
  float* sq = const_cast<float*>(A.a.sq) + (long long)blockIdx.x * A.stat_stride;
  for (int j = threadIdx.y; j < 16; j += PY) {
    float s = 0.f;
    for (int y = 0; y < PY; ++y) s += red[j * MAXT + y * G + cg];
    (j < 8 ? sum : sq)[c0 + (j & 7)] = s;
  }
}


// =====================================================================================================
// TMA-staged variants (plain bf16 NHWC tensors, no NCHW side input / output, no split pairs).
//
// The register-staged kernels above keep at most two pixel groups per thread in flight (126 registers, 16 warps per
// SM): ncu shows them latency-bound at 0.4-0.56 of the HBM peak with the same ~20 us per launch for 50 and for 75 MB
// (profiles/r02_bn_ncu.md).  Here the bytes in flight do not live in registers: every input tensor is streamed
// through a ring of shared-memory tiles by cp.async.bulk (1-D TMA, a tile = `tp` whole pixel rows = tp * Cp * 2
// contiguous bytes) completing on mbarriers; one thread issues the copy of tile t + S - 1 while the block consumes
// tile t (128-bit conflict-free shared-memory reads, results stored straight to global memory), one block barrier per
// tile hands the slot back.  A block owns a contiguous range of tiles of one problem.
constexpr int TS_MAX_IN = 4;
struct TileStreams {
  const __nv_bfloat16* src[TS_MAX_IN];
  int n;
};

__device__ __forceinline__ void bulk_load(void* dst, const void* src, uint32_t bytes, uint64_t* bar) {
  asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
               ::"r"(tc::smem_u32(dst)), "l"(src), "r"(bytes), "r"(tc::smem_u32(bar))
               : "memory");
}

struct TileRing {
  uint64_t* full;
  uint8_t* buf;
  int stages, tile_bytes, tp, Cp;
  long long M, t_lo, t_hi;
  __device__ __forceinline__ void init(uint8_t* sm, int stages_, int tp_, const mp_bn_args& A) {
    full = reinterpret_cast<uint64_t*>(sm);
    buf = sm + 128;
    stages = stages_; tp = tp_; Cp = A.Cp; M = A.M;
    tile_bytes = tp * Cp * 2;
    const long long tiles = (M + tp - 1) / tp;
    t_lo = (long long)blockIdx.x * tiles / gridDim.x;
    t_hi = (long long)(blockIdx.x + 1) * tiles / gridDim.x;
    if (threadIdx.x == 0 && threadIdx.y == 0) {
      for (int i = 0; i < stages; ++i) tc::mbar_init(&full[i], 1);
      tc::mbar_fence_init();
    }
  }
  __device__ __forceinline__ int npix(long long t) const {
    const long long left = M - t * tp;
    return left < tp ? (int)left : tp;
  }
  __device__ __forceinline__ void issue(long long t, const TileStreams& S) {   // one thread
    const int s = (int)((t - t_lo) % stages);
    const uint32_t bytes = (uint32_t)npix(t) * (uint32_t)Cp * 2u;
    tc::mbar_arrive_expect_tx(&full[s], bytes * (uint32_t)S.n);
    for (int i = 0; i < S.n; ++i)
      bulk_load(buf + (size_t)(s * S.n + i) * tile_bytes, S.src[i] + t * tp * Cp, bytes, &full[s]);
  }
  __device__ __forceinline__ const __nv_bfloat16* tile(long long t, int n_in, int i) const {
    const int s = (int)((t - t_lo) % stages);
    return reinterpret_cast<const __nv_bfloat16*>(buf + (size_t)(s * n_in + i) * tile_bytes);
  }
  __device__ __forceinline__ void wait(long long t) {
    const long long i = t - t_lo;
    tc::mbar_wait(&full[(int)(i % stages)], (uint32_t)((i / stages) & 1));
  }
};

__device__ __forceinline__ void lds8(const __nv_bfloat16* p, float (&v)[8]) {
  unpack8(*reinterpret_cast<const uint4*>(p), v);
}

__global__ void __launch_bounds__(MAXT, 2) bn_fwd_tma_kernel(const __grid_constant__ BnGroup GRP, int tp, int stages) {
  extern __shared__ __align__(128) uint8_t ts_smem[];
  const mp_bn_args& A = GRP.a[blockIdx.y];
  TileRing R;
  R.init(ts_smem, stages, tp, A);
  pdl_trigger();
  __syncthreads();
  pdl_wait();
  const int c0 = threadIdx.x * 8;
  const bool has_b = A.b.y != nullptr;
  const bool lead = threadIdx.x == 0 && threadIdx.y == 0;
  TileStreams S;
  S.n = (has_b || A.res) ? 2 : 1;
  S.src[0] = reinterpret_cast<const __nv_bfloat16*>(A.a.y);
  S.src[1] = reinterpret_cast<const __nv_bfloat16*>(has_b ? A.b.y : A.res);
  if (lead)
    for (long long t = R.t_lo; t < R.t_hi && t < R.t_lo + stages - 1; ++t) R.issue(t, S);
  float sa[8], ha[8], sb[8], hb[8];
  affine(A, A.a, c0, false, sa, ha);
  if (has_b) affine(A, A.b, c0, false, sb, hb);
  if (blockIdx.x == 0 && threadIdx.y == 0) {   // bookkeeping: saved statistics + running buffers
    MP_BN_BOOKKEEPING(A, A.a, c0);
    if (has_b) MP_BN_BOOKKEEPING(A, A.b, c0);
  }
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(A.out);
  for (long long t = R.t_lo; t < R.t_hi; ++t) {
    if (lead && t + stages - 1 < R.t_hi) R.issue(t + stages - 1, S);
    R.wait(t);
    const __nv_bfloat16* ta = R.tile(t, S.n, 0);
    const __nv_bfloat16* tb = R.tile(t, S.n, 1);
    const int np = R.npix(t);
    for (int p = threadIdx.y; p < np; p += blockDim.y) {
      float z[8], u[8];
      lds8(ta + p * A.Cp + c0, z);
#pragma unroll
      for (int j = 0; j < 8; ++j) z[j] = fmaf(z[j], sa[j], ha[j]);
      if (A.relu_a) {
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] = fmaxf(z[j], 0.f);
      }
      if (has_b) {
        lds8(tb + p * A.Cp + c0, u);
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] += fmaf(u[j], sb[j], hb[j]);
      } else if (A.res) {
        lds8(tb + p * A.Cp + c0, u);
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] += u[j];
      }
      if (A.relu_out) {
#pragma unroll
        for (int j = 0; j < 8; ++j) z[j] = fmaxf(z[j], 0.f);
      }
      if (c0 + 8 > A.C) {
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (c0 + j >= A.C) z[j] = 0.f;
      }
      *reinterpret_cast<uint4*>(out + (t * tp + p) * A.Cp + c0) = pack8(z);
    }
    __syncthreads();   // the slot of tile t is free for tile t + stages (issued at the top of the next iteration)
  }
}

// Streams of the backward kernels: 0 = dout, 1 = y_a, then y_b (second branch) and the forward output (post-ReLU mask).
__device__ __forceinline__ void bwd_streams(const mp_bn_args& A, TileStreams& S, int& i_yb, int& i_out) {
  S.n = 2;
  S.src[0] = reinterpret_cast<const __nv_bfloat16*>(A.dout);
  S.src[1] = reinterpret_cast<const __nv_bfloat16*>(A.a.y);
  i_yb = i_out = -1;
  if (A.b.y) { i_yb = S.n; S.src[S.n++] = reinterpret_cast<const __nv_bfloat16*>(A.b.y); }
  if (A.relu_out) { i_out = S.n; S.src[S.n++] = reinterpret_cast<const __nv_bfloat16*>(A.out); }
}

// dz of both branches at one pixel from the shared-memory tiles (same masks as grads_at)
__device__ __forceinline__ void grads_smem(const mp_bn_args& A, const __nv_bfloat16* tg, const __nv_bfloat16* tya,
                                           const __nv_bfloat16* tout, int off, int c0, const float (&sa)[8],
                                           const float (&ha)[8], float (&dza)[8], float (&dzb)[8], float (&ya)[8]) {
  float g[8];
  lds8(tg + off, g);
  if (tout) {
    float o[8];
    lds8(tout + off, o);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (!(o[j] > 0.f)) g[j] = 0.f;
  }
  lds8(tya + off, ya);
  if (c0 + 8 > A.C) {
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (c0 + j >= A.C) g[j] = 0.f;
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    dzb[j] = g[j];
    dza[j] = (A.relu_a && !(fmaf(ya[j], sa[j], ha[j]) > 0.f)) ? 0.f : g[j];
  }
}

__global__ void __launch_bounds__(MAXT, 2) bn_bwd_reduce_tma_kernel(const __grid_constant__ BnGroup GRP, int tp, int stages) {
  extern __shared__ __align__(128) uint8_t ts_smem[];
  const mp_bn_args& A = GRP.a[blockIdx.y];
  TileRing R;
  R.init(ts_smem, stages, tp, A);
  pdl_trigger();
  __syncthreads();
  pdl_wait();
  const int cg = threadIdx.x, c0 = cg * 8;
  const bool has_b = A.b.y != nullptr;
  const bool lead = threadIdx.x == 0 && threadIdx.y == 0;
  TileStreams S;
  int i_yb, i_out;
  bwd_streams(A, S, i_yb, i_out);
  if (lead)
    for (long long t = R.t_lo; t < R.t_hi && t < R.t_lo + stages - 1; ++t) R.issue(t, S);
  float sa[8], ha[8];
  affine(A, A.a, c0, true, sa, ha);
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  for (long long t = R.t_lo; t < R.t_hi; ++t) {
    if (lead && t + stages - 1 < R.t_hi) R.issue(t + stages - 1, S);
    R.wait(t);
    const __nv_bfloat16* tg = R.tile(t, S.n, 0);
    const __nv_bfloat16* tya = R.tile(t, S.n, 1);
    const __nv_bfloat16* tyb = has_b ? R.tile(t, S.n, i_yb) : nullptr;
    const __nv_bfloat16* tout = i_out >= 0 ? R.tile(t, S.n, i_out) : nullptr;
    const int np = R.npix(t);
    for (int p = threadIdx.y; p < np; p += blockDim.y) {
      const int off = p * A.Cp + c0;
      float dza[8], dzb[8], ya[8], yb[8];
      grads_smem(A, tg, tya, tout, off, c0, sa, ha, dza, dzb, ya);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        acc[j] += dza[j];
        acc[8 + j] = fmaf(dza[j], ya[j], acc[8 + j]);
      }
      if (has_b) {
        lds8(tyb + off, yb);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          acc[16 + j] += dzb[j];
          acc[24 + j] = fmaf(dzb[j], yb[j], acc[24 + j]);
        }
      }
    }
    __syncthreads();
  }
  // block reduction over the pixel rows in the (now idle) tile ring, then one atomic per (sum, channel)
  float* red = reinterpret_cast<float*>(R.buf);
  const int G = blockDim.x, PY = blockDim.y;
  const int tid = threadIdx.y * G + cg;
  const int nsum = has_b ? 32 : 16;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (j < nsum) red[j * MAXT + tid] = acc[j];
  __syncthreads();
  float* dst = A.sums + (long long)(blockIdx.x % A.stat_replicas) * 4 * A.Cp;
  for (int j = threadIdx.y; j < nsum; j += PY) {
    float s = 0.f;
    for (int y = 0; y < PY; ++y) s += red[j * MAXT + y * G + cg];
    atomicAdd(dst + (j >> 3) * A.Cp + c0 + (j & 7), s);
  }
}

__global__ void __launch_bounds__(MAXT, 2) bn_bwd_apply_tma_kernel(const __grid_constant__ BnGroup GRP, int tp, int stages) {
  extern __shared__ __align__(128) uint8_t ts_smem[];
  const mp_bn_args& A = GRP.a[blockIdx.y];
  TileRing R;
  R.init(ts_smem, stages, tp, A);
  pdl_trigger();
  __syncthreads();
  pdl_wait();
  const int c0 = threadIdx.x * 8;
  const bool has_b = A.b.y != nullptr;
  const bool lead = threadIdx.x == 0 && threadIdx.y == 0;
  TileStreams S;
  int i_yb, i_out;
  bwd_streams(A, S, i_yb, i_out);
  if (lead)
    for (long long t = R.t_lo; t < R.t_hi && t < R.t_lo + stages - 1; ++t) R.issue(t, S);
  const bool first = blockIdx.x == 0 && threadIdx.y == 0;
  float sa[8], ha[8], ka[8], ca[8], sb[8], kb[8], cb[8];
  affine(A, A.a, c0, true, sa, ha);
  float sa2[8];
  bwd_coefs(A, A.a, 0, c0, sa2, ka, ca, first);
  if (has_b) bwd_coefs(A, A.b, 1, c0, sb, kb, cb, first);
  __nv_bfloat16* dya = reinterpret_cast<__nv_bfloat16*>(A.a.dy);
  __nv_bfloat16* dyb = reinterpret_cast<__nv_bfloat16*>(A.b.dy);
  __nv_bfloat16* dres = reinterpret_cast<__nv_bfloat16*>(A.dres);
  for (long long t = R.t_lo; t < R.t_hi; ++t) {
    if (lead && t + stages - 1 < R.t_hi) R.issue(t + stages - 1, S);
    R.wait(t);
    const __nv_bfloat16* tg = R.tile(t, S.n, 0);
    const __nv_bfloat16* tya = R.tile(t, S.n, 1);
    const __nv_bfloat16* tyb = has_b ? R.tile(t, S.n, i_yb) : nullptr;
    const __nv_bfloat16* tout = i_out >= 0 ? R.tile(t, S.n, i_out) : nullptr;
    const int np = R.npix(t);
    for (int p = threadIdx.y; p < np; p += blockDim.y) {
      const int off = p * A.Cp + c0;
      const long long goff = (t * tp + p) * A.Cp + c0;
      float dza[8], dzb[8], ya[8], o[8];
      grads_smem(A, tg, tya, tout, off, c0, sa, ha, dza, dzb, ya);
#pragma unroll
      for (int j = 0; j < 8; ++j) o[j] = fmaf(sa2[j], dza[j], fmaf(ka[j], ya[j], ca[j]));
      if (dya) *reinterpret_cast<uint4*>(dya + goff) = pack8(o);
      if (has_b) {
        float yb[8];
        lds8(tyb + off, yb);
#pragma unroll
        for (int j = 0; j < 8; ++j) o[j] = fmaf(sb[j], dzb[j], fmaf(kb[j], yb[j], cb[j]));
        if (dyb) *reinterpret_cast<uint4*>(dyb + goff) = pack8(o);
      } else if (dres) {
        *reinterpret_cast<uint4*>(dres + goff) = pack8(dzb);
      }
    }
    __syncthreads();
  }
}

// Tile size / ring depth of the TMA-staged kernels: ~4 pixels per thread and tile (~16 KB per input tensor; half that
// with three or four inputs), as many stages as fit ~100 KB (two blocks per SM).
bool tma_plan(const mp_bn_args* a, int which, int n_in, dim3 block, int* tp, int* stages, size_t* smem) {
  if (!(g_bn_tma & which) || a->lo_delta != 0 || a->out_nchw || a->dout_nchw) return false;
  const int per_thread = n_in >= 3 ? 2 : 4;
  *tp = (int)block.y * per_thread;
  const size_t tile = (size_t)*tp * a->Cp * 2;
  int s = (int)((100 * 1024) / (tile * n_in));
  if (s > 8) s = 8;
  if (s < 2) return false;
  *stages = s;
  size_t need = 128 + (size_t)s * n_in * tile;
  const size_t red = 128 + 32 * MAXT * sizeof(float);   // the reduce kernel reuses the ring for its block reduction
  *smem = need > red ? need : red;
  return true;
}

// Eval-mode BatchNorm as a per-channel affine (folded into the producing conv's epilogue, igemm.cu): one block
// per BatchNorm of the network.
__global__ void bn_fold_eval_kernel(const mp_bn_fold_entry* __restrict__ table) {
  pdl_trigger();
  pdl_wait();
  const mp_bn_fold_entry E = table[blockIdx.x];
  for (int c = threadIdx.x; c < E.Cp; c += blockDim.x) {
    float scale = 0.f, shift = 0.f;
    if (c < E.C) {
      scale = E.gamma[c] * rsqrtf(E.running_var[c] + E.eps);
      const float mean = E.running_mean[c] - (E.conv_bias ? E.conv_bias[c] : 0.f);
      shift = fmaf(-mean, scale, E.beta[c]);
    }
    E.scale[c] = scale;
    E.shift[c] = shift;
  }
}

int check_args(const mp_bn_args* a, const char* what, bool bwd) {
  MP_CHECK_ARG(a, "%s: null args", what);
  MP_CHECK_ARG(a->a.y && a->M > 0 && a->C > 0 && a->Cp >= a->C && a->Cp % 8 == 0 && a->Cp / 8 <= MAXT,
               "%s: bad shape (M %lld, C %d, Cp %d)", what, (long long)a->M, a->C, a->Cp);
  MP_CHECK_ARG(a->a.gamma && a->a.beta, "%s: missing affine parameters", what);
  MP_CHECK_ARG(!a->b.y || (a->b.gamma && a->b.beta), "%s: missing affine parameters of branch b", what);
  MP_CHECK_ARG(!(a->b.y && a->res), "%s: a second BN branch and an identity residual are exclusive", what);
  MP_CHECK_ARG(a->stat_replicas >= 1 && a->stat_replicas <= 64, "%s: stat_replicas %d out of range", what,
               a->stat_replicas);
  if (!bwd) {
    MP_CHECK_ARG(a->out || a->out_nchw, "%s: no output", what);
    if (a->training) {
      MP_CHECK_ARG(a->a.sum && a->a.sq && (!a->b.y || (a->b.sum && a->b.sq)), "%s: training needs batch sums", what);
    } else {
      MP_CHECK_ARG(a->a.running_mean && a->a.running_var &&
                       (!a->b.y || (a->b.running_mean && a->b.running_var)),
                   "%s: eval needs running statistics", what);
    }
  } else {
    MP_CHECK_ARG(a->training, "%s: backward through eval-mode BatchNorm is not on the hot path", what);
    MP_CHECK_ARG((a->dout != nullptr) != (a->dout_nchw != nullptr), "%s: exactly one of dout / dout_nchw", what);
    MP_CHECK_ARG(a->sums && a->a.save_mean && a->a.save_invstd, "%s: missing saved statistics / workspace", what);
    MP_CHECK_ARG(!a->b.y || (a->b.save_mean && a->b.save_invstd), "%s: missing saved statistics of branch b", what);
    MP_CHECK_ARG(!a->relu_out || a->out, "%s: relu_out needs the forward output", what);
    MP_CHECK_ARG(!a->bwd_counter || (a->stat_replicas == 1 && a->a.coef && (!a->b.y || a->b.coef)),
                 "%s: the fused coefficient finalize needs coef buffers and un-replicated sums", what);
  }
  MP_CHECK_ARG(!(a->out_nchw || a->dout_nchw) || a->HW > 0, "%s: HW missing", what);
  MP_CHECK_ARG(a->lo_delta >= 0 && a->lo_delta % 8 == 0, "%s: lo_delta must be a non-negative multiple of 8", what);
  return MP_OK;
}

// Grid-stride kernels: at most `cap` blocks (2 per SM resident), so the per-block prologue and, in
// the reduce pass, the per-block atomics are amortised over many pixels.
void launch_dims(const mp_bn_args* a, dim3* grid, dim3* block, int u, int cap) {
  const int G = a->Cp / 8;
  int py = MAXT / G;
  if (py < 1) py = 1;
  *block = dim3(G, py);
  const long long ppb = (long long)py * u;
  long long blocks = (a->M + ppb - 1) / ppb;
  if (blocks > cap) blocks = cap;
  *grid = dim3((unsigned)blocks);
}

}  // namespace

static bool same_shape(const mp_bn_args* x, const mp_bn_args* y) {
  return x->M == y->M && x->C == y->C && x->Cp == y->Cp && x->HW == y->HW && x->training == y->training &&
         x->relu_a == y->relu_a && x->relu_out == y->relu_out && x->stat_replicas == y->stat_replicas &&
         x->lo_delta == y->lo_delta &&
         (x->b.y == nullptr) == (y->b.y == nullptr) && (x->res == nullptr) == (y->res == nullptr) &&
         (x->dout == nullptr) == (y->dout == nullptr) && (x->out_nchw == nullptr) == (y->out_nchw == nullptr);
}

static int make_group(const mp_bn_args* args, int n, const char* what, bool bwd, BnGroup* g) {
  MP_CHECK_ARG(args, "%s: null args", what);
  MP_CHECK_ARG(n >= 1 && n <= MP_MAX_GROUP, "%s: %d problems (1..%d)", what, n, MP_MAX_GROUP);
  for (int i = 0; i < n; ++i) {
    int rc = check_args(&args[i], what, bwd);
    if (rc != MP_OK) return rc;
    MP_CHECK_ARG(i == 0 || same_shape(&args[0], &args[i]), "%s: problem %d differs in shape", what, i);
  }
  for (int i = 0; i < MP_MAX_GROUP; ++i) g->a[i] = args[i < n ? i : 0];
  return MP_OK;
}

extern "C" {

int mp_bn_fwd(const mp_bn_args* a, void* stream) { return mp_bn_fwd_grouped(a, 1, stream); }
int mp_bn_bwd_reduce(const mp_bn_args* a, void* stream) { return mp_bn_bwd_reduce_grouped(a, 1, stream); }
int mp_bn_bwd_apply(const mp_bn_args* a, void* stream) { return mp_bn_bwd_apply_grouped(a, 1, stream); }

int mp_bn_fold_eval(const mp_bn_fold_entry* table, int n, void* stream) {
  MP_CHECK_ARG(table && n > 0, "mp_bn_fold_eval: bad arguments");
  MP_CUDA(mp_launch(bn_fold_eval_kernel, dim3((unsigned)n), dim3(64), 0, (cudaStream_t)stream, table));
  MP_CHECK_LAUNCH("mp_bn_fold_eval");
  return MP_OK;
}

int mp_bn_fwd_grouped(const mp_bn_args* args, int n, void* stream) {
  BnGroup g;
  int rc = make_group(args, n, "mp_bn_fwd", false, &g);
  if (rc != MP_OK) return rc;
  dim3 grid, block;
  launch_dims(args, &grid, &block, U, 2 * 148 / n);
  grid.y = n;
  {
    int tp, stages;
    size_t smem;
    if (args->out && tma_plan(args, 1, (args->b.y || args->res) ? 2 : 1, block, &tp, &stages, &smem)) {
      static bool attr = false;
      if (!attr) {
        MP_CUDA(cudaFuncSetAttribute(bn_fwd_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
        attr = true;
      }
      const long long tiles = (args->M + tp - 1) / tp;
      if (grid.x > tiles) grid.x = (unsigned)tiles;
      MP_CUDA(mp_launch(bn_fwd_tma_kernel, grid, block, smem, (cudaStream_t)stream, g, tp, stages));
      MP_CHECK_LAUNCH("mp_bn_fwd");
      return MP_OK;
    }
  }
  if (args->lo_delta) MP_CUDA(mp_launch(bn_fwd_kernel<true, 2>, grid, block, 0, (cudaStream_t)stream, g));
  else MP_CUDA(mp_launch(bn_fwd_kernel<false, 2>, grid, block, 0, (cudaStream_t)stream, g));
  MP_CHECK_LAUNCH("mp_bn_fwd");
  return MP_OK;
}

int mp_bn_stats_grouped(const mp_bn_args* args, int n, void* stream) {
  MP_CHECK_ARG(args, "mp_bn_stats: null args");
  MP_CHECK_ARG(n >= 1 && n <= MP_MAX_GROUP, "mp_bn_stats: %d problems (1..%d)", n, MP_MAX_GROUP);
  BnGroup g;
  for (int i = 0; i < n; ++i) {
    const mp_bn_args* a = &args[i];
    MP_CHECK_ARG(a->a.y && a->a.sum && a->a.sq && a->M > 0 && a->Cp % 8 == 0 && a->Cp / 8 <= MAXT && a->stat_replicas >= 1 &&
                     a->stat_replicas <= 64 && (a->stat_replicas == 1 || a->stat_stride >= a->Cp),
                 "mp_bn_stats: bad arguments");
    MP_CHECK_ARG(i == 0 || (a->M == args[0].M && a->Cp == args[0].Cp && a->stat_replicas == args[0].stat_replicas &&
                            a->lo_delta == args[0].lo_delta), "mp_bn_stats: problem %d differs in shape", i);
  }
  for (int i = 0; i < MP_MAX_GROUP; ++i) g.a[i] = args[i < n ? i : 0];
  const int G = args->Cp / 8;
  int py = MAXT / G;
  if (py < 1) py = 1;
  dim3 grid((unsigned)args->stat_replicas, (unsigned)n), block(G, py);
  if (args->lo_delta) MP_CUDA(mp_launch(bn_stats_kernel<true>, grid, block, 0, (cudaStream_t)stream, g));
  else MP_CUDA(mp_launch(bn_stats_kernel<false>, grid, block, 0, (cudaStream_t)stream, g));
  MP_CHECK_LAUNCH("mp_bn_stats");
  return MP_OK;
}
int mp_bn_stats(const mp_bn_args* a, void* stream) { return mp_bn_stats_grouped(a, 1, stream); }

int mp_bn_bwd_reduce_grouped(const mp_bn_args* args, int n, void* stream) {
  BnGroup g;
  int rc = make_group(args, n, "mp_bn_bwd_reduce", true, &g);
  if (rc != MP_OK) return rc;
  dim3 grid, block;
  int cap = 2 * 148 / n;   // two resident blocks per SM (one block per SM reads at ~3 TB/s, tools/bn_scale.py)
  // deterministic mode: one block per replica of the sums, so no two blocks ever add into the same address
  if (mp_deterministic() && cap > args->stat_replicas) cap = args->stat_replicas;
  launch_dims(args, &grid, &block, U, cap);
  grid.y = n;
  {
    int tp, stages;
    size_t smem;
    const int n_in = 2 + (args->b.y ? 1 : 0) + (args->relu_out ? 1 : 0);
    if (args->dout && tma_plan(args, 2, n_in, block, &tp, &stages, &smem)) {
      static bool attr = false;
      if (!attr) {
        MP_CUDA(cudaFuncSetAttribute(bn_bwd_reduce_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
        attr = true;
      }
      const long long tiles = (args->M + tp - 1) / tp;
      if (grid.x > tiles) grid.x = (unsigned)tiles;
      MP_CUDA(mp_launch(bn_bwd_reduce_tma_kernel, grid, block, smem, (cudaStream_t)stream, g, tp, stages));
      MP_CHECK_LAUNCH("mp_bn_bwd_reduce");
      return MP_OK;
    }
  }
  const bool split = args->lo_delta != 0;
  if (args->dout && split) MP_CUDA(mp_launch(bn_bwd_reduce_kernel<false, true, 2>, grid, block, 0, (cudaStream_t)stream, g));
  else if (args->dout) MP_CUDA(mp_launch(bn_bwd_reduce_kernel<false, false, 2>, grid, block, 0, (cudaStream_t)stream, g));
  else if (split) MP_CUDA(mp_launch(bn_bwd_reduce_kernel<true, true, 2>, grid, block, 0, (cudaStream_t)stream, g));
  else MP_CUDA(mp_launch(bn_bwd_reduce_kernel<true, false, 2>, grid, block, 0, (cudaStream_t)stream, g));
  MP_CHECK_LAUNCH("mp_bn_bwd_reduce");
  return MP_OK;
}

int mp_bn_bwd_apply_grouped(const mp_bn_args* args, int n, void* stream) {
  BnGroup g;
  int rc = make_group(args, n, "mp_bn_bwd_apply", true, &g);
  if (rc != MP_OK) return rc;
  dim3 grid, block;
  launch_dims(args, &grid, &block, UA, 2 * 148 / n);
  grid.y = n;
  {
    int tp, stages;
    size_t smem;
    const int n_in = 2 + (args->b.y ? 1 : 0) + (args->relu_out ? 1 : 0);
    if (args->dout && !args->bwd_counter && tma_plan(args, 4, n_in, block, &tp, &stages, &smem)) {
      static bool attr = false;
      if (!attr) {
        MP_CUDA(cudaFuncSetAttribute(bn_bwd_apply_tma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 112 * 1024));
        attr = true;
      }
      const long long tiles = (args->M + tp - 1) / tp;
      if (grid.x > tiles) grid.x = (unsigned)tiles;
      MP_CUDA(mp_launch(bn_bwd_apply_tma_kernel, grid, block, smem, (cudaStream_t)stream, g, tp, stages));
      MP_CHECK_LAUNCH("mp_bn_bwd_apply");
      return MP_OK;
    }
  }
  const bool split = args->lo_delta != 0;
  if (args->dout && split) MP_CUDA(mp_launch(bn_bwd_apply_kernel<false, true, 2>, grid, block, 0, (cudaStream_t)stream, g));
  else if (args->dout) MP_CUDA(mp_launch(bn_bwd_apply_kernel<false, false, 2>, grid, block, 0, (cudaStream_t)stream, g));
  else if (split) MP_CUDA(mp_launch(bn_bwd_apply_kernel<true, true, 2>, grid, block, 0, (cudaStream_t)stream, g));
  else MP_CUDA(mp_launch(bn_bwd_apply_kernel<true, false, 2>, grid, block, 0, (cudaStream_t)stream, g));
  MP_CHECK_LAUNCH("mp_bn_bwd_apply");
  return MP_OK;
}

}  // extern "C"
