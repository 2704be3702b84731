// BatchNorm + ReLU + residual, forward and backward, on bf16 NHWC activations (sm_100a).
//
// Replaces nn.BatchNorm2d / nn.ReLU / the residual `+` of ResidualBlock
// (/root/reference/src/margipose/models/margipose_model.py:31-40) and of the torchvision ResNet
// blocks (:130-135), and their autograd.  HBM-bound elementwise work: every tensor is touched once
// per pass with 128-bit loads/stores (8 bf16 channels per thread); the batch statistics arrive
// pre-reduced from the conv epilogue (igemm.cu), so the forward is a single pass.  Algorithmic
// bytes per pixel-channel: forward 2 per input tensor + 2 out; backward reduce 2 per tensor read,
// apply 2 per tensor read + 2 per gradient written.
#include "common.cuh"
#include "../../include/margipose_b200.h"

namespace {

constexpr int MAXT = 256;

struct Chan8 {
  float v[8];
};

__device__ __forceinline__ Chan8 load8(const __nv_bfloat16* p) {
  const uint4 u = __ldg(reinterpret_cast<const uint4*>(p));
  Chan8 r;
  float2 f;
  f = unpack_bf16x2(u.x); r.v[0] = f.x; r.v[1] = f.y;
  f = unpack_bf16x2(u.y); r.v[2] = f.x; r.v[3] = f.y;
  f = unpack_bf16x2(u.z); r.v[4] = f.x; r.v[5] = f.y;
  f = unpack_bf16x2(u.w); r.v[6] = f.x; r.v[7] = f.y;
  return r;
}
__device__ __forceinline__ void store8(__nv_bfloat16* p, const Chan8& r) {
  *reinterpret_cast<uint4*>(p) = make_uint4(pack_bf16x2(r.v[0], r.v[1]), pack_bf16x2(r.v[2], r.v[3]),
                                            pack_bf16x2(r.v[4], r.v[5]), pack_bf16x2(r.v[6], r.v[7]));
}

// Per-thread BatchNorm coefficients of 8 consecutive channels.
struct Coef {
  float scale[8], shift[8], mean[8], invstd[8];
};

// Forward: statistics from the conv epilogue sums (training) or the running buffers (eval).
__device__ __forceinline__ void coef_fwd(const mp_bn_branch& br, int c0, int C, long long M, int training,
                                         float eps, Coef& k) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + i;
    float mean = 0.f, invstd = 0.f, g = 0.f, b = 0.f;
    if (c < C) {
      if (training) {
        const float inv_m = 1.0f / (float)M;
        mean = br.sum[c] * inv_m;
        const float var = fmaxf(br.sq[c] * inv_m - mean * mean, 0.f);
        invstd = rsqrtf(var + eps);
      } else {
        mean = br.running_mean[c] - (br.conv_bias ? br.conv_bias[c] : 0.f);
        invstd = rsqrtf(br.running_var[c] + eps);
      }
      g = br.gamma[c];
      b = br.beta[c];
    }
    k.mean[i] = mean;
    k.invstd[i] = invstd;
    k.scale[i] = g * invstd;
    k.shift[i] = fmaf(-mean, k.scale[i], b);
  }
}
// Backward: the same coefficients, from the saved mean / invstd (bitwise equal to the forward's).
__device__ __forceinline__ void coef_bwd(const mp_bn_branch& br, int c0, int C, Coef& k) {
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int c = c0 + i;
    float mean = 0.f, invstd = 0.f, g = 0.f, b = 0.f;
    if (c < C) {
      mean = br.save_mean[c];
      invstd = br.save_invstd[c];
      g = br.gamma[c];
      b = br.beta[c];
    }
    k.mean[i] = mean;
    k.invstd[i] = invstd;
    k.scale[i] = g * invstd;
    k.shift[i] = fmaf(-mean, k.scale[i], b);
  }
}

// blockDim = (Cp/8 channel groups, PY pixel rows); each block walks `ppb` consecutive pixels.
__global__ void __launch_bounds__(MAXT) bn_fwd_kernel(const mp_bn_args A, int ppb) {
  const int cg = threadIdx.x, c0 = cg * 8;
  const bool has_b = A.b.y != nullptr;
  Coef ka, kb;
  coef_fwd(A.a, c0, A.C, A.M, A.training, A.eps, ka);
  if (has_b) coef_fwd(A.b, c0, A.C, A.M, A.training, A.eps, kb);

  if (blockIdx.x == 0 && threadIdx.y == 0) {   // bookkeeping: saved statistics + running buffers
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      const int c = c0 + i;
      if (c >= A.C) continue;
      for (int which = 0; which < (has_b ? 2 : 1); ++which) {
        const mp_bn_branch& br = which ? A.b : A.a;
        const Coef& k = which ? kb : ka;
        if (br.save_mean) {
          br.save_mean[c] = k.mean[i];
          br.save_invstd[c] = k.invstd[i];
        }
        if (A.training && br.running_mean) {
          const float inv_m = 1.0f / (float)A.M;
          const float var = fmaxf(br.sq[c] * inv_m - k.mean[i] * k.mean[i], 0.f);
          const float unbiased = A.M > 1 ? var * ((float)A.M / (float)(A.M - 1)) : var;
          const float bias = br.conv_bias ? br.conv_bias[c] : 0.f;
          br.running_mean[c] = (1.f - A.momentum) * br.running_mean[c] + A.momentum * (k.mean[i] + bias);
          br.running_var[c] = (1.f - A.momentum) * br.running_var[c] + A.momentum * unbiased;
        }
      }
    }
  }

  const __nv_bfloat16* ya = reinterpret_cast<const __nv_bfloat16*>(A.a.y);
  const __nv_bfloat16* yb = reinterpret_cast<const __nv_bfloat16*>(A.b.y);
  const __nv_bfloat16* res = reinterpret_cast<const __nv_bfloat16*>(A.res);
  __nv_bfloat16* out = reinterpret_cast<__nv_bfloat16*>(A.out);
  const long long p0 = (long long)blockIdx.x * ppb;
  for (int i = threadIdx.y; i < ppb; i += blockDim.y) {
    const long long pix = p0 + i;
    if (pix >= A.M) break;
    const long long off = pix * A.Cp + c0;
    Chan8 z = load8(ya + off);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      z.v[j] = fmaf(z.v[j], ka.scale[j], ka.shift[j]);
      if (A.relu_a) z.v[j] = fmaxf(z.v[j], 0.f);
    }
    if (has_b) {
      const Chan8 t = load8(yb + off);
#pragma unroll
      for (int j = 0; j < 8; ++j) z.v[j] += fmaf(t.v[j], kb.scale[j], kb.shift[j]);
    } else if (res) {
      const Chan8 t = load8(res + off);
#pragma unroll
      for (int j = 0; j < 8; ++j) z.v[j] += t.v[j];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (A.relu_out) z.v[j] = fmaxf(z.v[j], 0.f);
      if (c0 + j >= A.C) z.v[j] = 0.f;
    }
    if (out) store8(out + off, z);
    if (A.out_nchw) {
      const long long n = pix / A.HW, hw = pix - n * A.HW;
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (c0 + j < A.C) A.out_nchw[(n * A.C + c0 + j) * A.HW + hw] = z.v[j];
    }
  }
}

// dz of both branches at one pixel (shared by the reduce and apply passes).
struct Dz {
  Chan8 a, b, xa, xb;   // gradients w.r.t. the BN outputs and the normalised inputs x^
};

__device__ __forceinline__ void compute_dz(const mp_bn_args& A, const Coef& ka, const Coef& kb, bool has_b,
                                           long long pix, int c0, Dz& d) {
  const long long off = pix * A.Cp + c0;
  Chan8 g;
  if (A.dout) {
    g = load8(reinterpret_cast<const __nv_bfloat16*>(A.dout) + off);
  } else {
    const long long n = pix / A.HW, hw = pix - n * A.HW;
#pragma unroll
    for (int j = 0; j < 8; ++j)
      g.v[j] = (c0 + j < A.C) ? __ldg(A.dout_nchw + (n * A.C + c0 + j) * A.HW + hw) : 0.f;
  }
  if (A.relu_out) {
    const Chan8 o = load8(reinterpret_cast<const __nv_bfloat16*>(A.out) + off);
#pragma unroll
    for (int j = 0; j < 8; ++j)
      if (!(o.v[j] > 0.f)) g.v[j] = 0.f;
  }
  const Chan8 ya = load8(reinterpret_cast<const __nv_bfloat16*>(A.a.y) + off);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    float dz = g.v[j];
    if (A.relu_a && !(fmaf(ya.v[j], ka.scale[j], ka.shift[j]) > 0.f)) dz = 0.f;
    if (c0 + j >= A.C) dz = 0.f;
    d.a.v[j] = dz;
    d.xa.v[j] = (ya.v[j] - ka.mean[j]) * ka.invstd[j];
  }
  if (has_b) {
    const Chan8 yb = load8(reinterpret_cast<const __nv_bfloat16*>(A.b.y) + off);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      d.b.v[j] = (c0 + j < A.C) ? g.v[j] : 0.f;
      d.xb.v[j] = (yb.v[j] - kb.mean[j]) * kb.invstd[j];
    }
  } else {
#pragma unroll
    for (int j = 0; j < 8; ++j) d.b.v[j] = (c0 + j < A.C) ? g.v[j] : 0.f;   // identity residual gradient
  }
}

__global__ void __launch_bounds__(MAXT) bn_bwd_reduce_kernel(const mp_bn_args A, int ppb) {
  __shared__ float red[MAXT * 32];
  const int cg = threadIdx.x, c0 = cg * 8;
  const bool has_b = A.b.y != nullptr;
  Coef ka, kb;
  coef_bwd(A.a, c0, A.C, ka);
  if (has_b) coef_bwd(A.b, c0, A.C, kb);
  float acc[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) acc[j] = 0.f;
  const long long p0 = (long long)blockIdx.x * ppb;
  for (int i = threadIdx.y; i < ppb; i += blockDim.y) {
    const long long pix = p0 + i;
    if (pix >= A.M) break;
    Dz d;
    compute_dz(A, ka, kb, has_b, pix, c0, d);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      acc[j] += d.a.v[j];
      acc[8 + j] += d.a.v[j] * d.xa.v[j];
      if (has_b) {
        acc[16 + j] += d.b.v[j];
        acc[24 + j] += d.b.v[j] * d.xb.v[j];
      }
    }
  }
  const int G = blockDim.x, tid = threadIdx.y * G + cg;
  const int nsum = has_b ? 32 : 16;
#pragma unroll
  for (int j = 0; j < 32; ++j)
    if (j < nsum) red[j * MAXT + tid] = acc[j];
  __syncthreads();
  if (threadIdx.y == 0) {
    for (int j = 0; j < nsum; ++j) {
      float s = 0.f;
      for (int y = 0; y < (int)blockDim.y; ++y) s += red[j * MAXT + y * G + cg];
      atomicAdd(A.sums + (j >> 3) * A.Cp + c0 + (j & 7), s);
    }
  }
}

__global__ void __launch_bounds__(MAXT) bn_bwd_apply_kernel(const mp_bn_args A, int ppb) {
  const int cg = threadIdx.x, c0 = cg * 8;
  const bool has_b = A.b.y != nullptr;
  Coef ka, kb;
  coef_bwd(A.a, c0, A.C, ka);
  if (has_b) coef_bwd(A.b, c0, A.C, kb);
  const float inv_m = 1.0f / (float)A.M;
  float m1a[8], m2a[8], m1b[8], m2b[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    m1a[j] = A.sums[0 * A.Cp + c0 + j] * inv_m;
    m2a[j] = A.sums[1 * A.Cp + c0 + j] * inv_m;
    m1b[j] = has_b ? A.sums[2 * A.Cp + c0 + j] * inv_m : 0.f;
    m2b[j] = has_b ? A.sums[3 * A.Cp + c0 + j] * inv_m : 0.f;
  }
  if (blockIdx.x == 0 && threadIdx.y == 0) {   // affine-parameter gradients
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      if (c >= A.C) continue;
      if (A.a.dbeta) A.a.dbeta[c] += A.sums[0 * A.Cp + c];
      if (A.a.dgamma) A.a.dgamma[c] += A.sums[1 * A.Cp + c];
      if (has_b && A.b.dbeta) A.b.dbeta[c] += A.sums[2 * A.Cp + c];
      if (has_b && A.b.dgamma) A.b.dgamma[c] += A.sums[3 * A.Cp + c];
    }
  }
  __nv_bfloat16* dya = reinterpret_cast<__nv_bfloat16*>(A.a.dy);
  __nv_bfloat16* dyb = reinterpret_cast<__nv_bfloat16*>(A.b.dy);
  __nv_bfloat16* dres = reinterpret_cast<__nv_bfloat16*>(A.dres);
  const long long p0 = (long long)blockIdx.x * ppb;
  for (int i = threadIdx.y; i < ppb; i += blockDim.y) {
    const long long pix = p0 + i;
    if (pix >= A.M) break;
    const long long off = pix * A.Cp + c0;
    Dz d;
    compute_dz(A, ka, kb, has_b, pix, c0, d);
    Chan8 o;
#pragma unroll
    for (int j = 0; j < 8; ++j) o.v[j] = ka.scale[j] * (d.a.v[j] - m1a[j] - d.xa.v[j] * m2a[j]);
    if (dya) store8(dya + off, o);
    if (has_b) {
#pragma unroll
      for (int j = 0; j < 8; ++j) o.v[j] = kb.scale[j] * (d.b.v[j] - m1b[j] - d.xb.v[j] * m2b[j]);
      if (dyb) store8(dyb + off, o);
    } else if (dres) {
      store8(dres + off, d.b);
    }
  }
}

int check_args(const mp_bn_args* a, const char* what, bool bwd) {
  MP_CHECK_ARG(a, "%s: null args", what);
  MP_CHECK_ARG(a->a.y && a->M > 0 && a->C > 0 && a->Cp >= a->C && a->Cp % 8 == 0 && a->Cp / 8 <= MAXT,
               "%s: bad shape (M %lld, C %d, Cp %d)", what, (long long)a->M, a->C, a->Cp);
  MP_CHECK_ARG(a->a.gamma && a->a.beta, "%s: missing affine parameters", what);
  MP_CHECK_ARG(!a->b.y || (a->b.gamma && a->b.beta), "%s: missing affine parameters of branch b", what);
  MP_CHECK_ARG(!(a->b.y && a->res), "%s: a second BN branch and an identity residual are exclusive", what);
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
  }
  MP_CHECK_ARG(!(a->out_nchw || a->dout_nchw) || a->HW > 0, "%s: HW missing", what);
  return MP_OK;
}

void launch_dims(const mp_bn_args* a, int pix_per_thread, dim3* grid, dim3* block, int* ppb) {
  const int G = a->Cp / 8;
  int py = MAXT / G;
  if (py < 1) py = 1;
  *block = dim3(G, py);
  *ppb = py * pix_per_thread;
  *grid = dim3((unsigned)((a->M + *ppb - 1) / *ppb));
}

}  // namespace

extern "C" {

int mp_bn_fwd(const mp_bn_args* a, void* stream) {
  int rc = check_args(a, "mp_bn_fwd", false);
  if (rc != MP_OK) return rc;
  dim3 grid, block;
  int ppb;
  launch_dims(a, 4, &grid, &block, &ppb);
  bn_fwd_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(*a, ppb);
  MP_CHECK_LAUNCH("mp_bn_fwd");
  return MP_OK;
}

int mp_bn_bwd_reduce(const mp_bn_args* a, void* stream) {
  int rc = check_args(a, "mp_bn_bwd_reduce", true);
  if (rc != MP_OK) return rc;
  dim3 grid, block;
  int ppb;
  launch_dims(a, 8, &grid, &block, &ppb);
  bn_bwd_reduce_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(*a, ppb);
  MP_CHECK_LAUNCH("mp_bn_bwd_reduce");
  return MP_OK;
}

int mp_bn_bwd_apply(const mp_bn_args* a, void* stream) {
  int rc = check_args(a, "mp_bn_bwd_apply", true);
  if (rc != MP_OK) return rc;
  dim3 grid, block;
  int ppb;
  launch_dims(a, 4, &grid, &block, &ppb);
  bn_bwd_apply_kernel<<<grid, block, 0, (cudaStream_t)stream>>>(*a, ppb);
  MP_CHECK_LAUNCH("mp_bn_bwd_apply");
  return MP_OK;
}

}  // extern "C"
