// Fused differentiable tail of MargiPose for sm_100a: spatial softmax, soft-argmax (DSNT)
// expectations, xy/zy/xz marginal combination, Gaussian target rendering, Jensen-Shannon
// regulariser and Euclidean loss -- one forward kernel and one backward kernel.
//
// Replaces (reference, /root/reference/src/margipose/): dsntnn.py:124-130 (flat_softmax),
// :84-96 (dsnt), :154-195 (make_gauss), :198-232 (js_reg_losses), :133-151
// (euclidean_losses); models/margipose_model.py:223-261 (loss assembly, heatmaps_to_coords).
// Maths: SURVEY.md Appendix B.
//
// Bandwidth-bound: one CTA per (sample, joint) walks the three H x W planes; every heatmap
// element is read once (128-bit loads) and written once (128-bit stores); all reductions are
// warp shuffles + one shared-memory hop; no atomics, so results are run-to-run deterministic.
// Algorithmic HBM bytes per element: forward 8 (read logit, write prob), backward 12
// (read prob, read upstream grad, write dlogit).
#include "common.cuh"
#include "../../include/margipose_b200.h"

long long g_tail_fast = 1;   // tunable "tail_fast": 1 = the log2-domain kernels for row lengths dividing 128
long long g_tail_wpj_max = 32;   // tunable "tail_wpj_max": largest plane (in float4 per lane) of the forward warp-per-joint kernel
long long g_tail_cap = 4;    // tunable "tail_cap": float4 slots per lane the warp plan tries first (4 or 8)
long long g_tail_wpj = 2;    // tunable "tail_wpj": 1 = warp-per-joint kernels for planes of at most 1024 elements, 2 = also their multi-warp variants with the JS term / backward for larger planes
long long g_tail_waves = 0;  // tunable "tail_ctas_per_sm": persistent CTAs per SM of the fast kernels (0 = one CTA per group)

namespace {

constexpr int NT = 256;        // threads per CTA
constexpr int NW = NT / 32;
constexpr int MAX_DIM = 1024;  // max H or W
constexpr float KL_EPS = 1e-24f;

struct Geom {
  int H, W, HW;
  float cw_step, cw_first, ch_step, ch_first;  // pixel centre c_i = i * step + first
  float kw, kh;                                // Gaussian exponent scale per axis
};

struct FwdArgs {
  const float* in[3];
  float* prob[3];        // nullable
  float* ab[3];          // nullable: (BJ, 2) per-plane (col, row) expectations
  float* js[3];          // nullable: (BJ) per-plane JS divergence
  const float* target;   // nullable: (BJ, 3) xyz targets  (fused mode)
  const float* mu[3];    // nullable: (BJ, 2) per-plane (col,row) targets (generic mode)
  const int* valid_depth;  // nullable: (B) 1 = 3D sample, 0 = 2D sample
  float* coords;         // nullable: (BJ, 3)
  float* loss;           // nullable: (BJ)
  int J;
  int accumulate;        // loss[bj] += instead of =
  int pixelwise;         // 1 = JSD term on, 0 = off
  Geom g;
};

struct BwdArgs {
  const float* prob[3];
  const float* gup[3];    // nullable upstream grad on probs
  float* out[3];          // nullable -> plane skipped
  const float* target;    // fused mode: (BJ,3)
  const float* coords;    // fused mode: (BJ,3) saved by forward
  const float* w;         // fused mode: (BJ) dL/dloss[b,j]
  const int* valid_depth;
  const float* mu[3];     // generic mode: (BJ,2) per plane (needed when coef w_js != 0)
  const float* coef[3];   // generic mode: (BJ,3) = (w_js, c_col, c_row) per plane
  int J;
  int pixelwise;
  Geom g;
};

__device__ __forceinline__ float centre(int i, float step, float first) {
  return __fadd_rn(__fmul_rn((float)i, step), first);   // same op order as dsntnn.py:35-36
}

template <int N>
__device__ __forceinline__ void block_sum(float (&x)[N], float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
  for (int k = 0; k < N; ++k) x[k] = warp_sum(x[k]);
  if (lane == 0) {
#pragma unroll
    for (int k = 0; k < N; ++k) red[wid * N + k] = x[k];
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < N; ++k) {
    float s = 0.f;
#pragma unroll
    for (int w = 0; w < NW; ++w) s += red[w * N + k];   // fixed order -> deterministic
    x[k] = s;
  }
  __syncthreads();
}

__device__ __forceinline__ float block_max(float v, float* red) {
  const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
  v = warp_max(v);
  if (lane == 0) red[wid] = v;
  __syncthreads();
  float m = red[0];
#pragma unroll
  for (int w = 1; w < NW; ++w) m = fmaxf(m, red[w]);
  __syncthreads();
  return m;
}

template <int VEC, int V>
__device__ __forceinline__ void plane_load(const float* __restrict__ src, int HW, float fill,
                                           float (&x)[V * VEC]) {
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int e = (threadIdx.x + i * NT) * VEC;
    if (VEC == 4) {
      if (e < HW) {
        const float4 t = __ldg(reinterpret_cast<const float4*>(src + e));
        x[i * VEC + 0] = t.x; x[i * VEC + 1] = t.y; x[i * VEC + 2] = t.z; x[i * VEC + 3] = t.w;
      } else {
#pragma unroll
        for (int k = 0; k < VEC; ++k) x[i * VEC + k] = fill;
      }
    } else {
      x[i] = (e < HW) ? __ldg(src + e) : fill;
    }
  }
}

template <int VEC, int V>
__device__ __forceinline__ void plane_store(float* __restrict__ dst, int HW,
                                            const float (&x)[V * VEC]) {
#pragma unroll
  for (int i = 0; i < V; ++i) {
    const int e = (threadIdx.x + i * NT) * VEC;
    if (e < HW) {
      if (VEC == 4) {
        *reinterpret_cast<float4*>(dst + e) =
            make_float4(x[i * VEC + 0], x[i * VEC + 1], x[i * VEC + 2], x[i * VEC + 3]);
      } else {
        dst[e] = x[i];
      }
    }
  }
}

// Renders the separable Gaussian factors for one plane into shared memory and returns
// 1 / (sum + eps) (dsntnn.py:169-195).
__device__ __forceinline__ float gauss_factors(const Geom& g, float mu_col, float mu_row,
                                               float* s_ecol, float* s_erow, float* red) {
  float part[2] = {0.f, 0.f};
  for (int i = threadIdx.x; i < g.W; i += NT) {
    const float d = centre(i, g.cw_step, g.cw_first) - mu_col;
    const float e = expf(__fmul_rn(__fmul_rn(d, d), g.kw));
    s_ecol[i] = e;
    part[0] += e;
  }
  for (int i = threadIdx.x; i < g.H; i += NT) {
    const float d = centre(i, g.ch_step, g.ch_first) - mu_row;
    const float e = expf(__fmul_rn(__fmul_rn(d, d), g.kh));
    s_erow[i] = e;
    part[1] += e;
  }
  block_sum<2>(part, red);   // also publishes s_ecol / s_erow (it syncs)
  return 1.0f / (part[0] * part[1] + KL_EPS);
}

// One plane, forward. x: logits (FROM_LOGITS) or probabilities in; probabilities out.
// Results (a = col expectation, b = row expectation, js) are returned to every thread.
template <int VEC, int V, bool FROM_LOGITS>
__device__ __forceinline__ void plane_forward(float (&x)[V * VEC], const Geom& g, bool want_js,
                                              float mu_col, float mu_row, float* s_ecol,
                                              float* s_erow, float* red, float& a, float& b,
                                              float& js) {
  float ginv = 0.f;
  if (want_js) ginv = gauss_factors(g, mu_col, mu_row, s_ecol, s_erow, red);

  if (FROM_LOGITS) {
    float m = -INFINITY;
#pragma unroll
    for (int i = 0; i < V * VEC; ++i) m = fmaxf(m, x[i]);
    m = block_max(m, red);
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int e = (threadIdx.x + i * NT) * VEC;
      const int h = e / g.W, w0 = e - h * g.W;
      const float ch = centre(h, g.ch_step, g.ch_first);
#pragma unroll
      for (int k = 0; k < VEC; ++k) {
        const float ex = expf(x[i * VEC + k] - m);   // fill = -inf -> 0
        x[i * VEC + k] = ex;
        acc[0] += ex;
        acc[1] += ex * centre(w0 + k, g.cw_step, g.cw_first);
        acc[2] += ex * ch;
      }
    }
    block_sum<3>(acc, red);
    const float inv = 1.0f / acc[0];
#pragma unroll
    for (int i = 0; i < V * VEC; ++i) x[i] *= inv;
    a = acc[1] * inv;
    b = acc[2] * inv;
    js = 0.f;
    if (want_js) {
      float part[1] = {0.f};
#pragma unroll
      for (int i = 0; i < V; ++i) {
        const int e = (threadIdx.x + i * NT) * VEC;
        if (e < g.HW) {
          const int h = e / g.W, w0 = e - h * g.W;
          const float er = s_erow[h] * ginv;
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            const float p = x[i * VEC + k];
            const float q = s_ecol[w0 + k] * er;
            const float lm = logf(0.5f * (p + q) + KL_EPS);
            part[0] += p * (logf(p + KL_EPS) - lm) + q * (logf(q + KL_EPS) - lm);
          }
        }
      }
      block_sum<1>(part, red);
      js = 0.5f * part[0];
    }
  } else {
    float acc[3] = {0.f, 0.f, 0.f};
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int e = (threadIdx.x + i * NT) * VEC;
      if (e < g.HW) {
        const int h = e / g.W, w0 = e - h * g.W;
        const float ch = centre(h, g.ch_step, g.ch_first);
        const float er = want_js ? s_erow[h] * ginv : 0.f;
#pragma unroll
        for (int k = 0; k < VEC; ++k) {
          const float p = x[i * VEC + k];
          acc[0] += p * centre(w0 + k, g.cw_step, g.cw_first);
          acc[1] += p * ch;
          if (want_js) {
            const float q = s_ecol[w0 + k] * er;
            const float lm = logf(0.5f * (p + q) + KL_EPS);
            acc[2] += p * (logf(p + KL_EPS) - lm) + q * (logf(q + KL_EPS) - lm);
          }
        }
      }
    }
    block_sum<3>(acc, red);
    a = acc[0];
    b = acc[1];
    js = 0.5f * acc[2];
  }
}

template <int VEC, int V, bool FROM_LOGITS>
__global__ void __launch_bounds__(NT) tail_fwd_kernel(const FwdArgs A) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_ecol[MAX_DIM];
  __shared__ float s_erow[MAX_DIM];
  __shared__ float red[NW * 3];
  constexpr bool PRELOAD = (V * VEC <= 16);
  constexpr int NP = PRELOAD ? 3 : 1;
  float x[NP][V * VEC];

  const int bj = blockIdx.x;
  const int b = bj / A.J;
  const size_t off = (size_t)bj * A.g.HW;
  const float fill = FROM_LOGITS ? -INFINITY : 0.f;
  const bool is3d = A.valid_depth ? (A.valid_depth[b] != 0) : true;

  float tx = 0.f, ty = 0.f, tz = 0.f;
  if (A.target) {
    tx = A.target[bj * 3 + 0];
    ty = A.target[bj * 3 + 1];
    tz = A.target[bj * 3 + 2];
  }

  if (PRELOAD) {
#pragma unroll
    for (int k = 0; k < 3; ++k)
      if (A.in[k]) plane_load<VEC, V>(A.in[k] + off, A.g.HW, fill, x[PRELOAD ? k : 0]);
  }

  float ea[3] = {0.f, 0.f, 0.f}, eb[3] = {0.f, 0.f, 0.f}, js[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (!A.in[k]) continue;
    float(&xk)[V * VEC] = x[PRELOAD ? k : 0];
    if (!PRELOAD) plane_load<VEC, V>(A.in[k] + off, A.g.HW, fill, xk);
    // per-plane target (col, row): xy -> (x, y); zy -> (z, y); xz -> (x, z)
    float mc, mr;
    bool want_js;
    if (A.mu[k]) {
      mc = A.mu[k][bj * 2 + 0];
      mr = A.mu[k][bj * 2 + 1];
      want_js = A.js[k] != nullptr;
    } else {
      mc = (k == 1) ? tz : tx;
      mr = (k == 2) ? tz : ty;
      want_js = A.target && A.pixelwise && (A.loss || A.js[k]) && (k == 0 || is3d);
    }
    plane_forward<VEC, V, FROM_LOGITS>(xk, A.g, want_js, mc, mr, s_ecol, s_erow, red, ea[k], eb[k],
                                       js[k]);
    if (A.prob[k]) plane_store<VEC, V>(A.prob[k] + off, A.g.HW, xk);
    if (threadIdx.x == 0) {
      if (A.ab[k]) {
        A.ab[k][bj * 2 + 0] = ea[k];
        A.ab[k][bj * 2 + 1] = eb[k];
      }
      if (A.js[k]) A.js[k][bj] = js[k];
    }
  }

  if (threadIdx.x == 0) {
    // margipose_model.py:254-261
    const float px = ea[0], py = eb[0], pz = 0.5f * (ea[1] + eb[2]);
    if (A.coords) {
      A.coords[bj * 3 + 0] = px;
      A.coords[bj * 3 + 1] = py;
      A.coords[bj * 3 + 2] = pz;
    }
    if (A.loss && A.target) {
      const float dx = px - tx, dy = py - ty, dz = pz - tz;
      float l;
      if (is3d) l = js[0] + js[1] + js[2] + sqrtf(dx * dx + dy * dy + dz * dz);
      else l = js[0] + sqrtf(dx * dx + dy * dy);
      A.loss[bj] = A.accumulate ? A.loss[bj] + l : l;
    }
  }
}

// Backward of one (sample, joint): for each plane
//   D = G + w_js * dJS/dp + c_col * c_w + c_row * c_h ;  out = PROJECT ? p * (D - sum(p*D)) : D
template <int VEC, int V, bool PROJECT>
__global__ void __launch_bounds__(NT) tail_bwd_kernel(const BwdArgs A) {
  pdl_trigger();
  pdl_wait();
  __shared__ float s_ecol[MAX_DIM];
  __shared__ float s_erow[MAX_DIM];
  __shared__ float red[NW * 3];
  float p[V * VEC], d[V * VEC];

  const int bj = blockIdx.x;
  const int b = bj / A.J;
  const size_t off = (size_t)bj * A.g.HW;
  const bool is3d = A.valid_depth ? (A.valid_depth[b] != 0) : true;

  float tx = 0.f, ty = 0.f, tz = 0.f, w = 0.f, ddx = 0.f, ddy = 0.f, ddz = 0.f;
  if (A.target) {   // fused mode: derive the coefficients from saved coords / targets
    tx = A.target[bj * 3 + 0];
    ty = A.target[bj * 3 + 1];
    tz = A.target[bj * 3 + 2];
    w = A.w[bj];
    const float dx = A.coords[bj * 3 + 0] - tx, dy = A.coords[bj * 3 + 1] - ty;
    const float dz = is3d ? A.coords[bj * 3 + 2] - tz : 0.f;
    // d dist / d coord = diff / dist; infinite at dist == 0, like the reference (dsntnn.py:149-150)
    const float inv = w / sqrtf(dx * dx + dy * dy + dz * dz);
    ddx = dx * inv;
    ddy = dy * inv;
    ddz = dz * inv;
  }

#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (!A.out[k]) continue;
    float wjs, cc, cr, mc, mr;
    if (A.coef[k]) {
      wjs = A.coef[k][bj * 3 + 0];
      cc = A.coef[k][bj * 3 + 1];
      cr = A.coef[k][bj * 3 + 2];
      mc = A.mu[k] ? A.mu[k][bj * 2 + 0] : 0.f;
      mr = A.mu[k] ? A.mu[k][bj * 2 + 1] : 0.f;
    } else if (A.target) {
      mc = (k == 1) ? tz : tx;
      mr = (k == 2) ? tz : ty;
      wjs = (A.pixelwise && (k == 0 || is3d)) ? w : 0.f;
      cc = (k == 0) ? ddx : (k == 1 ? 0.5f * ddz : 0.f);
      cr = (k == 0) ? ddy : (k == 2 ? 0.5f * ddz : 0.f);
    } else {
      wjs = cc = cr = mc = mr = 0.f;
    }
    const bool want_js = (A.coef[k] ? (A.mu[k] != nullptr) : (wjs != 0.f));
    float ginv = 0.f;
    if (want_js) ginv = gauss_factors(A.g, mc, mr, s_ecol, s_erow, red);

    plane_load<VEC, V>(A.prob[k] + off, A.g.HW, 0.f, p);
    if (A.gup[k]) {
      plane_load<VEC, V>(A.gup[k] + off, A.g.HW, 0.f, d);
    } else {
#pragma unroll
      for (int i = 0; i < V * VEC; ++i) d[i] = 0.f;
    }
    float part[1] = {0.f};
#pragma unroll
    for (int i = 0; i < V; ++i) {
      const int e = (threadIdx.x + i * NT) * VEC;
      if (e < A.g.HW) {
        const int h = e / A.g.W, w0 = e - h * A.g.W;
        const float lin_r = cr * centre(h, A.g.ch_step, A.g.ch_first);
        const float er = want_js ? s_erow[h] * ginv : 0.f;
#pragma unroll
        for (int kk = 0; kk < VEC; ++kk) {
          const float pv = p[i * VEC + kk];
          float dv = d[i * VEC + kk] + cc * centre(w0 + kk, A.g.cw_step, A.g.cw_first) + lin_r;
          if (want_js) {
            const float q = s_ecol[w0 + kk] * er;
            const float m = 0.5f * (pv + q);
            const float djs = 0.5f * (logf(pv + KL_EPS) - logf(m + KL_EPS) + pv / (pv + KL_EPS) -
                                      m / (m + KL_EPS));
            dv += wjs * djs;
          }
          d[i * VEC + kk] = dv;
          part[0] += pv * dv;
        }
      }
    }
    if (PROJECT) {
      block_sum<1>(part, red);
#pragma unroll
      for (int i = 0; i < V * VEC; ++i) d[i] = p[i] * (d[i] - part[0]);
    }
    plane_store<VEC, V>(A.out[k] + off, A.g.HW, d);
    if (!PROJECT && want_js) __syncthreads();   // s_ecol / s_erow are rewritten by the next plane
  }
}

// ---- masked mean (dsntnn.py:99-121): out[0] = sum(l*m)/max(sum(m),1), out[1] = that denominator
__global__ void __launch_bounds__(NT) masked_mean_kernel(const float* __restrict__ l,
                                                         const float* __restrict__ m, int n,
                                                         float* __restrict__ out) {
  __shared__ float red[NW * 2];
  float acc[2] = {0.f, 0.f};
  for (int i = threadIdx.x; i < n; i += NT) {
    const float mv = m ? m[i] : 1.f;
    acc[0] += l[i] * mv;
    acc[1] += mv;
  }
  block_sum<2>(acc, red);
  if (threadIdx.x == 0) {
    const float den = fmaxf(acc[1], 1.f);
    out[0] = acc[0] / den;
    out[1] = den;
  }
}

// grad_l[i] = g[0] * m[i] / den
__global__ void masked_mean_bwd_kernel(const float* __restrict__ g, const float* __restrict__ m,
                                       const float* __restrict__ mean_den, int n,
                                       float* __restrict__ gl) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) gl[i] = g[0] * (m ? m[i] : 1.f) / mean_den[1];
}

// euclidean_losses (dsntnn.py:133-151): out[i] = |a[i,:] - t[i,:]|_2
__global__ void euclid_fwd_kernel(const float* __restrict__ a, const float* __restrict__ t, int n,
                                  int d, float* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  float s = 0.f;
  for (int k = 0; k < d; ++k) {
    const float v = a[i * d + k] - t[i * d + k];
    s += v * v;
  }
  out[i] = sqrtf(s);
}

// grad_a[i,k] = g[i] * (a - t)[i,k] / dist[i]   (infinite at dist == 0, as in the reference)
__global__ void euclid_bwd_kernel(const float* __restrict__ g, const float* __restrict__ a,
                                  const float* __restrict__ t, const float* __restrict__ dist,
                                  int n, int d, float* __restrict__ ga) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float s = g[i] / dist[i];
  for (int k = 0; k < d; ++k) ga[i * d + k] = (a[i * d + k] - t[i * d + k]) * s;
}

// make_gauss (dsntnn.py:154-195), 2D: out[bj, h, w]
__global__ void __launch_bounds__(NT) make_gauss_kernel(const float* __restrict__ mu,
                                                        float* __restrict__ out, Geom g,
                                                        int normalize) {
  __shared__ float s_ecol[MAX_DIM];
  __shared__ float s_erow[MAX_DIM];
  __shared__ float red[NW * 3];
  const int bj = blockIdx.x;
  float ginv = gauss_factors(g, mu[bj * 2 + 0], mu[bj * 2 + 1], s_ecol, s_erow, red);
  if (!normalize) ginv = 1.f;
  float* dst = out + (size_t)bj * g.HW;
  for (int e = threadIdx.x; e < g.HW; e += NT) {
    const int h = e / g.W, w = e - h * g.W;
    dst[e] = s_ecol[w] * (s_erow[h] * ginv);
  }
}

// =====================================================================================
// Warp-sliced fast path (16-byte aligned planes with W % 4 == 0 and H*W <= 4096*...): every
// plane is covered by `wpp` warps holding NV float4 per lane in registers, all loads are issued
// before the first use, reductions are warp shuffles plus ONE shared-memory hop per quantity, and a
// CTA runs `groups` (sample, joint) pairs x `np` planes side by side -- 4 block barriers per CTA
// instead of ~8 per plane, which is what lets the kernel approach the HBM roofline on 32x32 planes.
struct WarpPlan {
  int wpp, groups, np, BJ;
  int seq;      // 1: the CTA's warps all work on one plane at a time (large planes), planes in sequence
  int pid[3];   // plane slot -> plane index (only planes that are present get warps)
};

template <int NV, bool FROM_LOGITS>
__global__ void __launch_bounds__(512, 2) tail_fwd_warp_kernel(const FwdArgs A, const WarpPlan P) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  const int W = A.g.W, H = A.g.H, HW = A.g.HW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nslots = P.groups * P.np;
  const int nseq = P.seq ? P.np : 1;
  for (int it = 0; it < nseq; ++it) {
  const int slot = P.seq ? it : warp / P.wpp;
  const int s = P.seq ? warp : warp - slot * P.wpp;
  const int g = slot / P.np, ks = slot - g * P.np;
  const int k = P.pid[ks];
  const int bj = blockIdx.x * P.groups + g;
  const bool active = bj < P.BJ;
  float* tab = sm + slot * 2 * (W + H);                   // ecol[W], erow[H], then their logs
  float* ltab = tab + (W + H);
  float* red = sm + nslots * 2 * (W + H) + slot * (P.wpp * 4);   // per warp: sum, sum*cw, sum*ch, max
  float* jsr = sm + nslots * 2 * (W + H) + nslots * P.wpp * 4 + slot * P.wpp;   // per warp JS partial
  float* res = sm + nslots * 2 * (W + H) + nslots * P.wpp * 5 + slot * 2;       // per plane (a, b)

  const int b = active ? bj / A.J : 0;
  const bool is3d = (active && A.valid_depth) ? (A.valid_depth[b] != 0) : true;
  float tx = 0.f, ty = 0.f, tz = 0.f;
  if (active && A.target) {
    tx = A.target[bj * 3 + 0]; ty = A.target[bj * 3 + 1]; tz = A.target[bj * 3 + 2];
  }
  float mc = 0.f, mr = 0.f;
  bool want_js = false;
  if (active) {
    if (A.mu[k]) {
      mc = A.mu[k][bj * 2 + 0]; mr = A.mu[k][bj * 2 + 1];
      want_js = A.js[k] != nullptr;
    } else {
      mc = (k == 1) ? tz : tx;
      mr = (k == 2) ? tz : ty;
      want_js = A.target && A.pixelwise && (A.loss || A.js[k]) && (k == 0 || is3d);
    }
  }
  const size_t off = (size_t)bj * HW;
  const float fill = FROM_LOGITS ? -INFINITY : 0.f;
  float4 x[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int e = ((s * NV + i) * 32 + lane) * 4;
    x[i] = (active && e < HW) ? __ldg(reinterpret_cast<const float4*>(A.in[k] + off + e))
                              : make_float4(fill, fill, fill, fill);
  }
  float ginv = 0.f;
  if (want_js) {   // every warp of the plane computes the (tiny) normaliser; slice 0 publishes the factors
    float sc = 0.f, sr = 0.f;
    for (int i = lane; i < W; i += 32) {
      const float d = centre(i, A.g.cw_step, A.g.cw_first) - mc;
      const float lg = __fmul_rn(__fmul_rn(d, d), A.g.kw);
      const float ev = expf(lg);
      sc += ev;
      if (s == 0) { tab[i] = ev; ltab[i] = lg; }
    }
    for (int i = lane; i < H; i += 32) {
      const float d = centre(i, A.g.ch_step, A.g.ch_first) - mr;
      const float lg = __fmul_rn(__fmul_rn(d, d), A.g.kh);
      const float ev = expf(lg);
      sr += ev;
      if (s == 0) { tab[W + i] = ev; ltab[W + i] = lg; }
    }
    ginv = 1.0f / (warp_sum(sc) * warp_sum(sr) + KL_EPS);
  }
  const float log_ginv = want_js ? logf(ginv) : 0.f;
  float m = -INFINITY;
  if (FROM_LOGITS) {
#pragma unroll
    for (int i = 0; i < NV; ++i) m = fmaxf(m, fmaxf(fmaxf(x[i].x, x[i].y), fmaxf(x[i].z, x[i].w)));
    m = warp_max(m);
    if (lane == 0) red[s * 4 + 3] = m;
  }
  __syncthreads();   // #1: Gaussian factors + per-warp maxima visible
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f;
  if (FROM_LOGITS) {
    for (int q = 0; q < P.wpp; ++q) m = fmaxf(m, red[q * 4 + 3]);
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int e = ((s * NV + i) * 32 + lane) * 4;
    const int h = e / W, w0 = e - h * W;
    const float ch = centre(h, A.g.ch_step, A.g.ch_first);
    float v[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
    float rowsum = 0.f;
    // logits path: keep t = x - max in the registers (log p = t - log(sum) needs no log call later)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      float ev = v[j];
      if (FROM_LOGITS) {
        v[j] = v[j] - m;          // fill = -inf stays -inf -> exp = 0
        ev = __expf(v[j]);
      }
      rowsum += ev;
      acc1 += ev * centre(w0 + j, A.g.cw_step, A.g.cw_first);
    }
    acc0 += rowsum;
    acc2 += rowsum * ch;
    x[i] = make_float4(v[0], v[1], v[2], v[3]);
  }
  acc0 = warp_sum(acc0); acc1 = warp_sum(acc1); acc2 = warp_sum(acc2);
  if (lane == 0) { red[s * 4 + 0] = acc0; red[s * 4 + 1] = acc1; red[s * 4 + 2] = acc2; }
  __syncthreads();   // #2
  acc0 = acc1 = acc2 = 0.f;
  for (int q = 0; q < P.wpp; ++q) { acc0 += red[q * 4 + 0]; acc1 += red[q * 4 + 1]; acc2 += red[q * 4 + 2]; }
  float ea, eb;
  float inv = 1.f, log_inv = 0.f;
  if (FROM_LOGITS) {
    inv = 1.0f / acc0;
    log_inv = -logf(acc0);
    ea = acc1 * inv; eb = acc2 * inv;
  } else {
    ea = acc1; eb = acc2;
  }
  // probabilities + JS.  log p and log q come in closed form (log-softmax; the Gaussian's exponent),
  // clamped like log(. + 1e-24) clamps them; only log((p+q)/2) needs the special-function unit.
  const float LOG_EPS = -55.262042231857096f;   // log(1e-24)
  float jsp = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int e = ((s * NV + i) * 32 + lane) * 4;
    const int h = e / W, w0 = e - h * W;
    float pv[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
    float lp[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      if (FROM_LOGITS) {
        lp[j] = fmaxf(pv[j] + log_inv, LOG_EPS);
        pv[j] = __expf(pv[j]) * inv;
      } else if (want_js) {
        lp[j] = __logf(pv[j] + KL_EPS);
      }
    }
    x[i] = make_float4(pv[0], pv[1], pv[2], pv[3]);
    if (want_js && e < HW) {
      const float er = tab[W + h] * ginv;
      const float lr = ltab[W + h] + log_ginv;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float q = tab[w0 + j] * er;
        const float lq = fmaxf(ltab[w0 + j] + lr, LOG_EPS);
        const float lm = __logf(0.5f * (pv[j] + q) + KL_EPS);
        jsp += pv[j] * (lp[j] - lm) + q * (lq - lm);
      }
    }
  }
  if (want_js) jsp = warp_sum(jsp);
  if (lane == 0) {
    jsr[s] = jsp;
    if (s == 0) { res[0] = ea; res[1] = eb; }
  }
  if (active && A.prob[k]) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int e = ((s * NV + i) * 32 + lane) * 4;
      if (e < HW) *reinterpret_cast<float4*>(A.prob[k] + off + e) = x[i];
    }
  }
  __syncthreads();   // #3: per-plane results complete
  if (!active || s != 0 || lane != 0) continue;
  float js = 0.f;
  for (int q = 0; q < P.wpp; ++q) js += jsr[q];
  js *= 0.5f;
  if (A.ab[k]) { A.ab[k][bj * 2 + 0] = ea; A.ab[k][bj * 2 + 1] = eb; }
  if (A.js[k]) A.js[k][bj] = js;
  if (ks != P.np - 1 && P.seq) continue;     // sequential mode: combine after the last plane
  if (ks != 0 && !P.seq) continue;
  // one leader combines the planes (models/margipose_model.py:254-261)
  float pa[3] = {0.f, 0.f, 0.f}, pb[3] = {0.f, 0.f, 0.f}, pj[3] = {0.f, 0.f, 0.f};
  for (int q = 0; q < P.np; ++q) {
    const int slot_q = g * P.np + q;
    const float* rq = sm + nslots * 2 * (W + H) + nslots * P.wpp * 5 + slot_q * 2;
    const float* jq = sm + nslots * 2 * (W + H) + nslots * P.wpp * 4 + slot_q * P.wpp;
    float t = 0.f;
    for (int u = 0; u < P.wpp; ++u) t += jq[u];
    pa[P.pid[q]] = rq[0]; pb[P.pid[q]] = rq[1]; pj[P.pid[q]] = 0.5f * t;
  }
  const float px = pa[0], py = pb[0], pz = 0.5f * (pa[1] + pb[2]);
  if (A.coords) { A.coords[bj * 3 + 0] = px; A.coords[bj * 3 + 1] = py; A.coords[bj * 3 + 2] = pz; }
  if (A.loss && A.target) {
    const float dx = px - tx, dy = py - ty, dz = pz - tz;
    float l;
    if (is3d) l = pj[0] + pj[1] + pj[2] + sqrtf(dx * dx + dy * dy + dz * dz);
    else l = pj[0] + sqrtf(dx * dx + dy * dy);
    A.loss[bj] = A.accumulate ? A.loss[bj] + l : l;
  }
  }
}

template <int NV, bool PROJECT>
__global__ void __launch_bounds__(512, 2) tail_bwd_warp_kernel(const BwdArgs A, const WarpPlan P) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  const int W = A.g.W, H = A.g.H, HW = A.g.HW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nslots = P.groups * P.np;
  const int nseq = P.seq ? P.np : 1;
  for (int it = 0; it < nseq; ++it) {
  const int slot = P.seq ? it : warp / P.wpp;
  const int s = P.seq ? warp : warp - slot * P.wpp;
  const int g = slot / P.np, ks = slot - g * P.np;
  const int k = P.pid[ks];
  const int bj = blockIdx.x * P.groups + g;
  const bool active = bj < P.BJ;
  float* tab = sm + slot * (W + H);
  float* red = sm + nslots * (W + H) + slot * P.wpp;

  float wjs = 0.f, cc = 0.f, cr = 0.f, mc = 0.f, mr = 0.f;
  bool want_js = false;
  if (active) {
    const int b = bj / A.J;
    const bool is3d = A.valid_depth ? (A.valid_depth[b] != 0) : true;
    if (A.coef[k]) {
      wjs = A.coef[k][bj * 3 + 0]; cc = A.coef[k][bj * 3 + 1]; cr = A.coef[k][bj * 3 + 2];
      mc = A.mu[k] ? A.mu[k][bj * 2 + 0] : 0.f;
      mr = A.mu[k] ? A.mu[k][bj * 2 + 1] : 0.f;
      want_js = A.mu[k] != nullptr;
    } else if (A.target) {
      const float tx = A.target[bj * 3 + 0], ty = A.target[bj * 3 + 1], tz = A.target[bj * 3 + 2];
      const float w = A.w[bj];
      const float dx = A.coords[bj * 3 + 0] - tx, dy = A.coords[bj * 3 + 1] - ty;
      const float dz = is3d ? A.coords[bj * 3 + 2] - tz : 0.f;
      const float inv = w / sqrtf(dx * dx + dy * dy + dz * dz);   // infinite at 0, like dsntnn.py:149-150
      mc = (k == 1) ? tz : tx;
      mr = (k == 2) ? tz : ty;
      wjs = (A.pixelwise && (k == 0 || is3d)) ? w : 0.f;
      cc = (k == 0) ? dx * inv : (k == 1 ? 0.5f * dz * inv : 0.f);
      cr = (k == 0) ? dy * inv : (k == 2 ? 0.5f * dz * inv : 0.f);
      want_js = wjs != 0.f;
    }
  }
  const size_t off = (size_t)bj * HW;
  float4 p[NV], d[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int e = ((s * NV + i) * 32 + lane) * 4;
    const bool ok = active && e < HW;
    p[i] = ok ? __ldg(reinterpret_cast<const float4*>(A.prob[k] + off + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
    d[i] = (ok && A.gup[k]) ? __ldg(reinterpret_cast<const float4*>(A.gup[k] + off + e))
                            : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  float ginv = 0.f;
  if (want_js) {
    float sc = 0.f, sr = 0.f;
    for (int i = lane; i < W; i += 32) {
      const float dd = centre(i, A.g.cw_step, A.g.cw_first) - mc;
      const float ev = expf(__fmul_rn(__fmul_rn(dd, dd), A.g.kw));
      sc += ev;
      if (s == 0) tab[i] = ev;
    }
    for (int i = lane; i < H; i += 32) {
      const float dd = centre(i, A.g.ch_step, A.g.ch_first) - mr;
      const float ev = expf(__fmul_rn(__fmul_rn(dd, dd), A.g.kh));
      sr += ev;
      if (s == 0) tab[W + i] = ev;
    }
    ginv = 1.0f / (warp_sum(sc) * warp_sum(sr) + KL_EPS);
  }
  __syncthreads();   // #1
  float part = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int e = ((s * NV + i) * 32 + lane) * 4;
    if (e < HW) {
      const int h = e / W, w0 = e - h * W;
      const float lin_r = cr * centre(h, A.g.ch_step, A.g.ch_first);
      const float er = want_js ? tab[W + h] * ginv : 0.f;
      const float pv[4] = {p[i].x, p[i].y, p[i].z, p[i].w};
      float dv[4] = {d[i].x, d[i].y, d[i].z, d[i].w};
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        dv[j] += cc * centre(w0 + j, A.g.cw_step, A.g.cw_first) + lin_r;
        if (want_js) {
          const float q = tab[w0 + j] * er;
          const float mm = 0.5f * (pv[j] + q);
          dv[j] += wjs * 0.5f * (__logf(pv[j] + KL_EPS) - __logf(mm + KL_EPS) +
                                 __fdividef(pv[j], pv[j] + KL_EPS) - __fdividef(mm, mm + KL_EPS));
        }
        part += pv[j] * dv[j];
      }
      d[i] = make_float4(dv[0], dv[1], dv[2], dv[3]);
    }
  }
  if (PROJECT) {
    part = warp_sum(part);
    if (lane == 0) red[s] = part;
    __syncthreads();   // #2
    part = 0.f;
    for (int q = 0; q < P.wpp; ++q) part += red[q];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      d[i].x = p[i].x * (d[i].x - part); d[i].y = p[i].y * (d[i].y - part);
      d[i].z = p[i].z * (d[i].z - part); d[i].w = p[i].w * (d[i].w - part);
    }
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int e = ((s * NV + i) * 32 + lane) * 4;
      if (e < HW) *reinterpret_cast<float4*>(A.out[k] + off + e) = d[i];
    }
  }
  if (!PROJECT && P.seq) __syncthreads();
  }
}


// =====================================================================================
// Fast variants of the warp-sliced kernels for row lengths that divide 128 (W = 32 / 64 / 128, the MargiPose
// heatmaps): a lane's four columns are the same in every one of its float4 slots, so the column centres and the
// Gaussian's column factors are per-lane CONSTANTS, rows advance by 128 / W per slot (no integer division), the
// softmax / JS arithmetic runs in the log2 domain (one FFMA + one MUFU.EX2 per exponential, one MUFU.LG2 per
// logarithm), and exp() is evaluated once per element unless the JS term needs log p.  ~7 (softmax + dsnt),
// ~18 (full forward) and ~16 (full backward) instructions per heatmap element instead of 48 / 83 / 83
// (profiles/r02_tail_ncu.md), which is what moves these kernels from issue-bound to HBM-bound.
constexpr float L2E = 1.4426950408889634f;    // log2(e)
constexpr float LN2 = 0.6931471805599453f;
// single MUFU instructions (no denormal fix-up code around them: exp2 of a very negative number is 0 either way,
// and every logarithm below is taken of a value >= 1e-24)
__device__ __forceinline__ float fast_ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float fast_lg2(float x) {
  float y;
  asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Sum / max of the `wpp` per-warp partials p[0], p[stride], ...: a short serial loop, or -- many warps per plane --
// one partial per lane and a shuffle tree (10 instead of ~4 * wpp instructions).  Both orders are fixed.
__device__ __forceinline__ float cross_warp_sum(const float* p, int stride, int wpp, int lane) {
  if (wpp <= 4) {
    float s = 0.f;
    for (int q = 0; q < wpp; ++q) s += p[q * stride];
    return s;
  }
  return warp_sum(lane < wpp ? p[lane * stride] : 0.f);
}
__device__ __forceinline__ float cross_warp_max(const float* p, int stride, int wpp, int lane, float m) {
  if (wpp <= 4) {
    for (int q = 0; q < wpp; ++q) m = fmaxf(m, p[q * stride]);
    return m;
  }
  return warp_max(lane < wpp ? p[lane * stride] : -INFINITY);
}

// Gaussian factor tables of one plane, computed ONCE by the plane's warps together (entry i by warp i / 32):
// tab[0..W) = column factors, tab[W..W+H) = row factors, ltab = their exponents (natural log), and per-warp partial
// sums for the normaliser in gs[warp * 3 + {0, 1}] (slot 2: the forward kernel's entropy partial).
__device__ __forceinline__ void gauss_tables(const Geom& g, float mc, float mr, int s, int wpp, int lane, float* tab,
                                             float* ltab, float* gs) {
  float sc = 0.f, sr = 0.f;
  for (int i = s * 32 + lane; i < g.W + g.H; i += wpp * 32) {
    const bool col = i < g.W;
    const float c = col ? centre(i, g.cw_step, g.cw_first) - mc : centre(i - g.W, g.ch_step, g.ch_first) - mr;
    const float lg = __fmul_rn(__fmul_rn(c, c), col ? g.kw : g.kh);
    const float ev = expf(lg);
    tab[i] = ev;
    if (ltab) ltab[i] = lg;
    if (col) sc += ev; else sr += ev;
  }
  sc = warp_sum(sc);
  sr = warp_sum(sr);
  if (lane == 0) { gs[s * 3] = sc; gs[s * 3 + 1] = sr; }
}

template <int NV, bool FROM_LOGITS>
__global__ void __launch_bounds__(512, 2) tail_fwd_fast_kernel(const FwdArgs A, const WarpPlan P, const int wshift) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  const int W = A.g.W, H = A.g.H, HW = A.g.HW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nslots = P.groups * P.np;
  const int nseq = P.seq ? P.np : 1;
  const int w0 = (lane * 4) & (W - 1);                 // this lane's four columns, in every slot
  const int rpi = 128 >> wshift;                       // rows a warp advances per float4 slot
  // persistent: a CTA walks (sample, joint) groups blockIdx.x, blockIdx.x + gridDim.x, ... -- the slot / pointer
  // set-up of a warp does not depend on the group and is paid once
  const int ngroups = (P.BJ + P.groups - 1) / P.groups;
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {
  for (int it = 0; it < nseq; ++it) {
  const int slot = P.seq ? it : warp / P.wpp;
  const int s = P.seq ? warp : warp - slot * P.wpp;
  const int g = slot / P.np, ks = slot - g * P.np;
  const int k = P.pid[ks];
  const int bj = grp * P.groups + g;
  const bool active = bj < P.BJ;
  float* tab = sm + slot * 2 * (W + H);                   // ecol[W], erow[H], then their exponents
  float* ltab = tab + (W + H);
  float* red = sm + nslots * 2 * (W + H) + slot * (P.wpp * 4);   // per warp: sum, sum*cw, sum*ch, max
  float* jsr = sm + nslots * 2 * (W + H) + nslots * P.wpp * 4 + slot * P.wpp;   // per warp JS partial
  float* res = sm + nslots * 2 * (W + H) + nslots * P.wpp * 5 + slot * 2;       // per plane (a, b)
  float* gs = sm + nslots * 2 * (W + H) + nslots * P.wpp * 5 + nslots * 2 + slot * (P.wpp * 3);   // normaliser partials

  const int b = active ? bj / A.J : 0;
  const bool is3d = (active && A.valid_depth) ? (A.valid_depth[b] != 0) : true;
  float tx = 0.f, ty = 0.f, tz = 0.f;
  if (active && A.target) {
    tx = A.target[bj * 3 + 0]; ty = A.target[bj * 3 + 1]; tz = A.target[bj * 3 + 2];
  }
  float mc = 0.f, mr = 0.f;
  bool want_js = false;
  if (active) {
    if (A.mu[k]) {
      mc = A.mu[k][bj * 2 + 0]; mr = A.mu[k][bj * 2 + 1];
      want_js = A.js[k] != nullptr;
    } else {
      mc = (k == 1) ? tz : tx;
      mr = (k == 2) ? tz : ty;
      want_js = A.target && A.pixelwise && (A.loss || A.js[k]) && (k == 0 || is3d);
    }
  }
  const size_t off = (size_t)bj * HW;
  const int e0 = (s * NV * 32 + lane) * 4;               // slot i covers elements e0 + i * 128 .. + 3
  const int h0 = e0 >> wshift;
  // (a large finite fill instead of -inf keeps exp2(t) * t = 0 * finite in the entropy sum of padded lanes)
  const float fill = FROM_LOGITS ? -1e30f : 0.f;
  float4 x[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int e = e0 + i * 128;
    x[i] = (active && e < HW) ? __ldg(reinterpret_cast<const float4*>(A.in[k] + off + e))
                              : make_float4(fill, fill, fill, fill);
  }
  if (want_js) gauss_tables(A.g, mc, mr, s, P.wpp, lane, tab, ltab, gs);
  float m = -INFINITY;
  if (FROM_LOGITS) {
#pragma unroll
    for (int i = 0; i < NV; ++i) m = fmaxf(m, fmaxf(fmaxf(x[i].x, x[i].y), fmaxf(x[i].z, x[i].w)));
    m = warp_max(m);
    if (lane == 0) red[s * 4 + 3] = m;
  }
  __syncthreads();   // #1: Gaussian factors + per-warp maxima visible
  float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accx = 0.f;
  if (FROM_LOGITS) m = cross_warp_max(red + 3, 4, P.wpp, lane, m);
  const float m2 = m * L2E;
  {
    float cw[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) cw[j] = centre(w0 + j, A.g.cw_step, A.g.cw_first);
    // pass 1: exponentials (log2 domain), partition sum, the two coordinate moments and -- for the JS term --
    // sum exp(t) * t (the entropy of p in closed form: log2 p = t2 - log2(sum)).  The registers then keep
    // t2 = (x - max) * log2(e) when a JS term follows (pass 2 re-exponentiates with the normaliser folded in),
    // otherwise exp(t).  (Two separate loops: a per-element select between the two would keep both alive.)
    if (FROM_LOGITS && want_js) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float v[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
        float rowsum = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[j] = fmaxf(fmaf(v[j], L2E, -m2), -1e30f);   // -inf logits stay finite: 0 * t2 = 0 below
          const float ev = fast_ex2(v[j]);
          accx = fmaf(ev, v[j], accx);
          rowsum += ev;
          acc1 = fmaf(ev, cw[j], acc1);
          v[j] = ev;
        }
        acc0 += rowsum;
        acc2 = fmaf(rowsum, centre(h0 + i * rpi, A.g.ch_step, A.g.ch_first), acc2);
        x[i] = make_float4(v[0], v[1], v[2], v[3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float v[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
        float rowsum = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (FROM_LOGITS) v[j] = fast_ex2(fmaf(v[j], L2E, -m2));
          rowsum += v[j];
          acc1 = fmaf(v[j], cw[j], acc1);
        }
        acc0 += rowsum;
        acc2 = fmaf(rowsum, centre(h0 + i * rpi, A.g.ch_step, A.g.ch_first), acc2);
        x[i] = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
  }
  acc0 = warp_sum(acc0); acc1 = warp_sum(acc1); acc2 = warp_sum(acc2);
  if (FROM_LOGITS && want_js) accx = warp_sum(accx);
  if (lane == 0) {
    red[s * 4 + 0] = acc0; red[s * 4 + 1] = acc1; red[s * 4 + 2] = acc2;
    if (FROM_LOGITS) gs[s * 3 + 2] = accx;
  }
  __syncthreads();   // #2
  accx = 0.f;
  acc0 = cross_warp_sum(red + 0, 4, P.wpp, lane);
  acc1 = cross_warp_sum(red + 1, 4, P.wpp, lane);
  acc2 = cross_warp_sum(red + 2, 4, P.wpp, lane);
  float ea, eb;
  float inv = 1.f, l2inv = 0.f;
  if (FROM_LOGITS) {
    inv = 1.0f / acc0;
    l2inv = -log2f(acc0);
    ea = acc1 * inv; eb = acc2 * inv;
  } else {
    ea = acc1; eb = acc2;
  }
  // pass 2: probabilities and JS in bits,
  //   2 JS / ln 2 = sum p lg p + sum q lg q - sum (p + q) lg((p + q) / 2 + eps)
  // sum p lg p comes from pass 1 (logits) and sum q lg q is separable in the Gaussian's row / column factors, so
  // an element costs one exp2, one log2 and five FP32 instructions.  (log(. + 1e-24) of the reference differs from
  // log(.) only where p or q < 1e-17, where p lg p and q lg q vanish in fp32 anyway.)
  float jsp = 0.f;
  if (want_js) {
    const float sc = cross_warp_sum(gs + 0, 3, P.wpp, lane), sr = cross_warp_sum(gs + 1, 3, P.wpp, lane);
    const float ginv = 1.0f / (sc * sr + KL_EPS);
    float qc[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) qc[j] = tab[w0 + j];
    float slm = 0.f, plp = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      float pv[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
      if (e0 + i * 128 < HW) {
        const float er = tab[W + h0 + i * rpi] * ginv;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (FROM_LOGITS) pv[j] *= inv;
          else plp = fmaf(pv[j], fast_lg2(pv[j] + KL_EPS), plp);
          const float sq = fmaf(qc[j], er, pv[j]);                 // p + q
          slm = fmaf(sq, fast_lg2(fmaf(0.5f, sq, KL_EPS)), slm);
        }
      } else if (FROM_LOGITS) {
        pv[0] = pv[1] = pv[2] = pv[3] = 0.f;
      }
      x[i] = make_float4(pv[0], pv[1], pv[2], pv[3]);
    }
    jsp = plp - slm;
    if (s == 0) {
      // this plane's sum q lg q = lg(ginv) + ginv * (sum_w qc lqc * sum_h qr + sum_w qc * sum_h qr lqr) * log2(e)
      // (sum q = 1 up to the 1e-24 in the normaliser), accumulated by warp 0 from the 1-D tables
      float a = 0.f, bq = 0.f;
      for (int i = lane; i < W; i += 32) a = fmaf(tab[i], ltab[i], a);
      for (int i = lane; i < H; i += 32) bq = fmaf(tab[W + i], ltab[W + i], bq);
      a = warp_sum(a); bq = warp_sum(bq);
      if (lane == 0) jsp += fmaf(ginv * L2E, fmaf(a, sr, sc * bq), log2f(ginv) * (sc * sr * ginv));
    }
    jsp = warp_sum(jsp);
    if (FROM_LOGITS && s == 0) {   // sum p lg p = (sum exp(t) t2) / sum exp(t) - log2(sum exp(t))
      jsp += fmaf(cross_warp_sum(gs + 2, 3, P.wpp, lane), inv, l2inv);
    }
    jsp *= LN2;
  } else if (FROM_LOGITS) {
#pragma unroll
    for (int i = 0; i < NV; ++i) { x[i].x *= inv; x[i].y *= inv; x[i].z *= inv; x[i].w *= inv; }
  }
  if (active && A.prob[k]) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int e = e0 + i * 128;
      if (e < HW) *reinterpret_cast<float4*>(A.prob[k] + off + e) = x[i];
    }
  }
  if (lane == 0) {
    jsr[s] = jsp;
    if (s == 0) { res[0] = ea; res[1] = eb; }
  }
  __syncthreads();   // #3: per-plane results complete
  if (!active || s != 0 || lane != 0) continue;
  float js = 0.f;
  for (int q = 0; q < P.wpp; ++q) js += jsr[q];
  js *= 0.5f;
  if (A.ab[k]) { A.ab[k][bj * 2 + 0] = ea; A.ab[k][bj * 2 + 1] = eb; }
  if (A.js[k]) A.js[k][bj] = js;
  if (ks != P.np - 1 && P.seq) continue;     // sequential mode: combine after the last plane
  if (ks != 0 && !P.seq) continue;
  // one leader combines the planes (models/margipose_model.py:254-261)
  float pa[3] = {0.f, 0.f, 0.f}, pb[3] = {0.f, 0.f, 0.f}, pj[3] = {0.f, 0.f, 0.f};
  for (int q = 0; q < P.np; ++q) {
    const int slot_q = g * P.np + q;
    const float* rq = sm + nslots * 2 * (W + H) + nslots * P.wpp * 5 + slot_q * 2;
    const float* jq = sm + nslots * 2 * (W + H) + nslots * P.wpp * 4 + slot_q * P.wpp;
    float t = 0.f;
    for (int u = 0; u < P.wpp; ++u) t += jq[u];
    pa[P.pid[q]] = rq[0]; pb[P.pid[q]] = rq[1]; pj[P.pid[q]] = 0.5f * t;
  }
  const float px = pa[0], py = pb[0], pz = 0.5f * (pa[1] + pb[2]);
  if (A.coords) { A.coords[bj * 3 + 0] = px; A.coords[bj * 3 + 1] = py; A.coords[bj * 3 + 2] = pz; }
  if (A.loss && A.target) {
    const float dx = px - tx, dy = py - ty, dz = pz - tz;
    float l;
    if (is3d) l = pj[0] + pj[1] + pj[2] + sqrtf(dx * dx + dy * dy + dz * dz);
    else l = pj[0] + sqrtf(dx * dx + dy * dy);
    A.loss[bj] = A.accumulate ? A.loss[bj] + l : l;
  }
  }
  __syncthreads();   // the next group rewrites the tables and partial sums the leaders may still be reading
  }
}

template <int NV, bool PROJECT>
__global__ void __launch_bounds__(512, 2) tail_bwd_fast_kernel(const BwdArgs A, const WarpPlan P, const int wshift) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  const int W = A.g.W, H = A.g.H, HW = A.g.HW;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nslots = P.groups * P.np;
  const int nseq = P.seq ? P.np : 1;
  const int w0 = (lane * 4) & (W - 1);
  const int rpi = 128 >> wshift;
  const int ngroups = (P.BJ + P.groups - 1) / P.groups;
  for (int grp = blockIdx.x; grp < ngroups; grp += gridDim.x) {   // persistent, see tail_fwd_fast_kernel
  for (int it = 0; it < nseq; ++it) {
  const int slot = P.seq ? it : warp / P.wpp;
  const int s = P.seq ? warp : warp - slot * P.wpp;
  const int g = slot / P.np, ks = slot - g * P.np;
  const int k = P.pid[ks];
  const int bj = grp * P.groups + g;
  const bool active = bj < P.BJ;
  float* tab = sm + slot * (W + H);
  float* red = sm + nslots * (W + H) + slot * P.wpp;
  float* gs = sm + nslots * (W + H) + nslots * P.wpp + slot * (P.wpp * 3);
  // large planes (8 float4 of p and of the gradient per lane would not fit 64 registers): the un-projected gradient
  // waits for the plane-wide sum in shared memory, one conflict-free float4 per lane and slot
  constexpr bool STASH = PROJECT && NV > 4;
  float4* stash = reinterpret_cast<float4*>(sm + ((nslots * (W + H) + nslots * P.wpp * 4 + 3) & ~3)) + threadIdx.x;

  float wjs = 0.f, cc = 0.f, cr = 0.f, mc = 0.f, mr = 0.f;
  bool want_js = false;
  if (active) {
    const int b = bj / A.J;
    const bool is3d = A.valid_depth ? (A.valid_depth[b] != 0) : true;
    if (A.coef[k]) {
      wjs = A.coef[k][bj * 3 + 0]; cc = A.coef[k][bj * 3 + 1]; cr = A.coef[k][bj * 3 + 2];
      mc = A.mu[k] ? A.mu[k][bj * 2 + 0] : 0.f;
      mr = A.mu[k] ? A.mu[k][bj * 2 + 1] : 0.f;
      want_js = A.mu[k] != nullptr;
    } else if (A.target) {
      const float tx = A.target[bj * 3 + 0], ty = A.target[bj * 3 + 1], tz = A.target[bj * 3 + 2];
      const float w = A.w[bj];
      const float dx = A.coords[bj * 3 + 0] - tx, dy = A.coords[bj * 3 + 1] - ty;
      const float dz = is3d ? A.coords[bj * 3 + 2] - tz : 0.f;
      const float inv = w / sqrtf(dx * dx + dy * dy + dz * dz);   // infinite at 0, like dsntnn.py:149-150
      mc = (k == 1) ? tz : tx;
      mr = (k == 2) ? tz : ty;
      wjs = (A.pixelwise && (k == 0 || is3d)) ? w : 0.f;
      cc = (k == 0) ? dx * inv : (k == 1 ? 0.5f * dz * inv : 0.f);
      cr = (k == 0) ? dy * inv : (k == 2 ? 0.5f * dz * inv : 0.f);
      want_js = wjs != 0.f;
    }
  }
  const size_t off = (size_t)bj * HW;
  const int e0 = (s * NV * 32 + lane) * 4;
  const int h0 = e0 >> wshift;
  float4 p[NV], d[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const int e = e0 + i * 128;
    const bool ok = active && e < HW;
    p[i] = ok ? __ldg(reinterpret_cast<const float4*>(A.prob[k] + off + e)) : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (!STASH) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int e = e0 + i * 128;
      d[i] = (active && e < HW && A.gup[k]) ? __ldg(reinterpret_cast<const float4*>(A.gup[k] + off + e))
                                            : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  }
  if (want_js) gauss_tables(A.g, mc, mr, s, P.wpp, lane, tab, nullptr, gs);
  __syncthreads();   // #1
  float ginv = 0.f;
  float ccw[4], qc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
  for (int j = 0; j < 4; ++j) ccw[j] = cc * centre(w0 + j, A.g.cw_step, A.g.cw_first);
  if (want_js) {
    const float sc = cross_warp_sum(gs + 0, 3, P.wpp, lane), sr = cross_warp_sum(gs + 1, 3, P.wpp, lane);
    ginv = 1.0f / (sc * sr + KL_EPS);
#pragma unroll
    for (int j = 0; j < 4; ++j) qc[j] = tab[w0 + j];
  }
  // dJS/dp = 0.5 (ln p' - ln m' + p / p' - m / m'), p' = p + eps, m' = (p + q) / 2 + eps.  eps = 1e-24 is invisible
  // to fp32 unless p < ~1e-16 (then m may be that small too): the fast form 0.5 ln 2 (lg p - lg(p + q) + 1) is used
  // for a float4 unless some lane of the warp sees such a p, in which case the warp evaluates the full expression.
  const float kjs = 0.5f * wjs * LN2, hjs = 0.5f * wjs;
  const float TINY = 2e-16f;
  float part = 0.f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    if (e0 + i * 128 < HW) {
      const int h = h0 + i * rpi;
      const float lin_r = cr * centre(h, A.g.ch_step, A.g.ch_first);
      const float pv[4] = {p[i].x, p[i].y, p[i].z, p[i].w};
      if (STASH)   // the upstream gradient is fetched two slots ahead of its use (L2 / HBM latency)
        d[i] = A.gup[k] ? __ldg(reinterpret_cast<const float4*>(A.gup[k] + off + e0 + i * 128))
                        : make_float4(0.f, 0.f, 0.f, 0.f);
      float dv[4] = {d[i].x, d[i].y, d[i].z, d[i].w};
      if (want_js) {
        const float er = tab[W + h] * ginv;
        const float lin_js = lin_r + kjs;                      // the "+ 1" of lg p - lg(p + q) + 1
        const bool tiny = fminf(fminf(pv[0], pv[1]), fminf(pv[2], pv[3])) < TINY;
        if (!__any_sync(0xffffffffu, tiny)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float sq = fmaf(qc[j], er, pv[j]);
            dv[j] += ccw[j] + lin_js;
            dv[j] = fmaf(kjs, fast_lg2(pv[j]) - fast_lg2(sq), dv[j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float mm = 0.5f * fmaf(qc[j], er, pv[j]);
            dv[j] += ccw[j] + lin_r;
            dv[j] = fmaf(kjs, fast_lg2(pv[j] + KL_EPS) - fast_lg2(mm + KL_EPS), dv[j]);
            dv[j] = fmaf(hjs, __fdividef(pv[j], pv[j] + KL_EPS) - __fdividef(mm, mm + KL_EPS), dv[j]);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) dv[j] += ccw[j] + lin_r;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) part = fmaf(pv[j], dv[j], part);
      if (STASH) stash[i * blockDim.x] = make_float4(dv[0], dv[1], dv[2], dv[3]);
      else d[i] = make_float4(dv[0], dv[1], dv[2], dv[3]);
    }
  }
  if (PROJECT) {
    part = warp_sum(part);
    if (lane == 0) red[s] = part;
    __syncthreads();   // #2
    part = cross_warp_sum(red, 1, P.wpp, lane);
    if (!STASH) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        d[i].x = p[i].x * (d[i].x - part); d[i].y = p[i].y * (d[i].y - part);
        d[i].z = p[i].z * (d[i].z - part); d[i].w = p[i].w * (d[i].w - part);
      }
    }
  }
  if (active) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const int e = e0 + i * 128;
      if (e < HW) {
        float4 o = d[i];
        if (STASH) {
          const float4 t = stash[i * blockDim.x];
          o = make_float4(p[i].x * (t.x - part), p[i].y * (t.y - part), p[i].z * (t.z - part), p[i].w * (t.w - part));
        }
        *reinterpret_cast<float4*>(A.out[k] + off + e) = o;
      }
    }
  }
  __syncthreads();   // the next plane / group rewrites the tables and partial sums
  }
  }
}


// =====================================================================================================
// Warp-per-joint kernels for small planes (H * W <= 1024 and a multiple of 128, W dividing 128: the 32 x 32 heatmaps
// of the MargiPose model itself, 16 x 16, ...).  ONE warp owns a (sample, joint) and walks its three planes, a whole
// plane lives in the warp's registers (NV float4 per lane), so there is no shared memory and no block barrier at all:
// every reduction is a shuffle tree, and the separable Gaussian's column / row factors are evaluated directly by the
// lanes that need them (4 columns + NV rows per lane) instead of going through a table.  The per-plane fixed cost
// drops from ~800 instructions per warp (fast kernels above) to ~200.
// WPP > 1 (planes above one warp's registers, up to 128 x 128): WPP warps = one block share a (sample, joint), each holds
// NV float4 per lane of the plane (forward: 32 without a JS term, 16 with it; backward: 8, it holds p and the gradient); their
// partial maxima / sums meet in shared memory (one block barrier per reduction, below).
//
// Block-wide sums / maximum of the WPP warps that share a plane: shuffle tree, lane 0 of every warp posts its N partial
// values as one float4, ONE block barrier, every warp folds the WPP posts with log2(WPP) more shuffles.  Consecutive
// reductions alternate between two shared-memory buffers (`phase`), so a buffer is rewritten only after a later barrier
// has proved that everyone finished reading it.
template <int WPP, int N>
__device__ __forceinline__ void wpj_sum_n(float (&v)[N], float4* red, int& phase, int warp, int lane) {
  static_assert(N <= 4, "one float4 per warp");
#pragma unroll
  for (int n = 0; n < N; ++n) v[n] = warp_sum(v[n]);
  if (WPP == 1) return;
  float4* r = red + (phase & 1) * 16;
  ++phase;
  if (lane == 0) r[warp] = make_float4(v[0], N > 1 ? v[N > 1 ? 1 : 0] : 0.f, N > 2 ? v[N > 2 ? 2 : 0] : 0.f, N > 3 ? v[N > 3 ? 3 : 0] : 0.f);
  __syncthreads();
  float4 t = r[lane & (WPP - 1)];
#pragma unroll
  for (int o = WPP >> 1; o > 0; o >>= 1) {
    t.x += __shfl_xor_sync(0xffffffffu, t.x, o);
    if (N > 1) t.y += __shfl_xor_sync(0xffffffffu, t.y, o);
    if (N > 2) t.z += __shfl_xor_sync(0xffffffffu, t.z, o);
    if (N > 3) t.w += __shfl_xor_sync(0xffffffffu, t.w, o);
  }
  v[0] = t.x;
  if (N > 1) v[N > 1 ? 1 : 0] = t.y;
  if (N > 2) v[N > 2 ? 2 : 0] = t.z;
  if (N > 3) v[N > 3 ? 3 : 0] = t.w;
}
template <int WPP>
__device__ __forceinline__ float wpj_sum(float v, float4* red, int& phase, int warp, int lane) {
  float a[1] = {v};
  wpj_sum_n<WPP, 1>(a, red, phase, warp, lane);
  return a[0];
}
template <int WPP>
__device__ __forceinline__ float wpj_max(float v, float4* red, int& phase, int warp, int lane) {
  v = warp_max(v);
  if (WPP == 1) return v;
  float4* r = red + (phase & 1) * 16;
  ++phase;
  if (lane == 0) r[warp].x = v;
  __syncthreads();
  float m = r[lane & (WPP - 1)].x;
#pragma unroll
  for (int o = WPP >> 1; o > 0; o >>= 1) m = fmaxf(m, __shfl_xor_sync(0xffffffffu, m, o));
  return m;
}

template <int NV, bool FROM_LOGITS, int WPP>
__global__ void __launch_bounds__(WPP > 1 ? 32 * WPP : (NV > 8 ? 128 : 256), (WPP > 1 && NV <= 16) ? (WPP >= 16 ? 1 : 16 / WPP) : 2)
tail_fwd_wpj_kernel(const FwdArgs A, const int BJ, const int wshift) {
  pdl_trigger();
  pdl_wait();
  __shared__ float4 red[WPP > 1 ? 32 : 1];
  int phase = 0;
  const int W = A.g.W, H = A.g.H, HW = A.g.HW;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sl = WPP > 1 ? warp : 0;             // this warp's slice of the plane (WPP > 1: one (sample, joint) per block)
  const int bj = WPP > 1 ? (int)blockIdx.x : blockIdx.x * (blockDim.x >> 5) + warp;
  if (bj >= BJ) return;
  const int w0 = (lane * 4) & (W - 1);
  const int rpi = 128 >> wshift;                 // rows between a lane's consecutive float4 slots
  const int h0 = ((sl * NV * 32 + lane) * 4) >> wshift;
  const float col_mult = (float)W * (1.0f / 128.0f);   // a column is held by 128 / W lanes, a row by W / 4
  const float row_mult = 4.0f / (float)W;
  float cw[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) cw[j] = centre(w0 + j, A.g.cw_step, A.g.cw_first);
  const int b = bj / A.J;
  const bool is3d = A.valid_depth ? (A.valid_depth[b] != 0) : true;
  float tx = 0.f, ty = 0.f, tz = 0.f;
  if (A.target) { tx = A.target[bj * 3 + 0]; ty = A.target[bj * 3 + 1]; tz = A.target[bj * 3 + 2]; }
  const size_t off = (size_t)bj * HW;
  float pa[3] = {0.f, 0.f, 0.f}, pb[3] = {0.f, 0.f, 0.f}, pj[3] = {0.f, 0.f, 0.f};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (!A.in[k]) continue;
    float4 x[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i)
      x[i] = __ldg(reinterpret_cast<const float4*>(A.in[k] + off + ((sl * NV + i) * 32 + lane) * 4));
    float mc, mr;
    bool want_js;
    if (A.mu[k]) {
      mc = A.mu[k][bj * 2 + 0]; mr = A.mu[k][bj * 2 + 1];
      want_js = A.js[k] != nullptr;
    } else {
      mc = (k == 1) ? tz : tx;
      mr = (k == 2) ? tz : ty;
      want_js = A.target && A.pixelwise && (A.loss || A.js[k]) && (k == 0 || is3d);
    }
    float m = 0.f;
    if (FROM_LOGITS) {
      m = -INFINITY;
#pragma unroll
      for (int i = 0; i < NV; ++i) m = fmaxf(m, fmaxf(fmaxf(x[i].x, x[i].y), fmaxf(x[i].z, x[i].w)));
      m = wpj_max<WPP>(m, red, phase, warp, lane);
    }
    const float m2 = m * L2E;
    float acc0 = 0.f, acc1 = 0.f, acc2 = 0.f, accx = 0.f;
    if (FROM_LOGITS && want_js) {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float v[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
        float rowsum = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          v[j] = fmaxf(fmaf(v[j], L2E, -m2), -1e30f);
          const float ev = fast_ex2(v[j]);
          accx = fmaf(ev, v[j], accx);
          rowsum += ev;
          acc1 = fmaf(ev, cw[j], acc1);
          v[j] = ev;
        }
        acc0 += rowsum;
        acc2 = fmaf(rowsum, centre(h0 + i * rpi, A.g.ch_step, A.g.ch_first), acc2);
        x[i] = make_float4(v[0], v[1], v[2], v[3]);
      }
    } else {
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float v[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
        float rowsum = 0.f;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (FROM_LOGITS) v[j] = fast_ex2(fmaf(v[j], L2E, -m2));
          rowsum += v[j];
          acc1 = fmaf(v[j], cw[j], acc1);
        }
        acc0 += rowsum;
        acc2 = fmaf(rowsum, centre(h0 + i * rpi, A.g.ch_step, A.g.ch_first), acc2);
        x[i] = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
    {
      float a3[3] = {acc0, acc1, acc2};
      wpj_sum_n<WPP, 3>(a3, red, phase, warp, lane);
      acc0 = a3[0]; acc1 = a3[1]; acc2 = a3[2];
    }
    float inv = 1.f, l2inv = 0.f, ea, eb;
    if (FROM_LOGITS) {
      inv = 1.0f / acc0;
      l2inv = -log2f(acc0);
      ea = acc1 * inv; eb = acc2 * inv;
    } else {
      ea = acc1; eb = acc2;
    }
    float js = 0.f;
    if (want_js) {
      // separable Gaussian, evaluated by the lanes that need it: 4 column factors, NV row factors (log2-domain exponents;
      // the row factors are re-evaluated in the JS pass -- one exp2 per float4 -- rather than kept in NV registers)
      float qc[4];
      float sc = 0.f, sr = 0.f, ac = 0.f, ar = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float d = cw[j] - mc;
        const float lg = __fmul_rn(__fmul_rn(d, d), A.g.kw) * L2E;
        qc[j] = fast_ex2(lg);
        sc += qc[j];
        ac = fmaf(qc[j], lg, ac);
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float d = centre(h0 + i * rpi, A.g.ch_step, A.g.ch_first) - mr;
        const float lg = __fmul_rn(__fmul_rn(d, d), A.g.kh) * L2E;
        const float qri = fast_ex2(lg);
        sr += qri;
        ar = fmaf(qri, lg, ar);
      }
      sc = warp_sum(sc) * col_mult; ac = warp_sum(ac) * col_mult;     // every warp holds whole rows: all columns
      {
        float r2[2] = {sr, ar};                                          // the rows are spread over the plane's warps
        wpj_sum_n<WPP, 2>(r2, red, phase, warp, lane);
        sr = r2[0] * row_mult; ar = r2[1] * row_mult;
      }
      const float ginv = 1.0f / (sc * sr + KL_EPS);
      float slm = 0.f, plp = 0.f;
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        float pv[4] = {x[i].x, x[i].y, x[i].z, x[i].w};
        const float dr = centre(h0 + i * rpi, A.g.ch_step, A.g.ch_first) - mr;
        const float er = fast_ex2(__fmul_rn(__fmul_rn(dr, dr), A.g.kh) * L2E) * ginv;
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          if (FROM_LOGITS) pv[j] *= inv;
          else plp = fmaf(pv[j], fast_lg2(pv[j] + KL_EPS), plp);
          const float sq = fmaf(qc[j], er, pv[j]);
          slm = fmaf(sq, fast_lg2(fmaf(0.5f, sq, KL_EPS)), slm);
        }
        x[i] = make_float4(pv[0], pv[1], pv[2], pv[3]);
      }
      float jsp = wpj_sum<WPP>(FROM_LOGITS ? fmaf(accx, inv, plp - slm) : plp - slm, red, phase, warp, lane);
      if (FROM_LOGITS) jsp += l2inv;                                   // sum p lg p = (sum e t) / (sum e) - lg2(sum e)
      jsp += fmaf(ginv, fmaf(ac, sr, sc * ar), log2f(ginv) * (sc * sr * ginv));   // sum q lg q (exponents are log2 already)
      js = 0.5f * jsp * LN2;
    } else if (FROM_LOGITS) {
#pragma unroll
      for (int i = 0; i < NV; ++i) { x[i].x *= inv; x[i].y *= inv; x[i].z *= inv; x[i].w *= inv; }
    }
    if (A.prob[k]) {
#pragma unroll
      for (int i = 0; i < NV; ++i)
        *reinterpret_cast<float4*>(A.prob[k] + off + ((sl * NV + i) * 32 + lane) * 4) = x[i];
    }
    pa[k] = ea; pb[k] = eb; pj[k] = js;
    if (lane == 0 && sl == 0) {
      if (A.ab[k]) { A.ab[k][bj * 2 + 0] = ea; A.ab[k][bj * 2 + 1] = eb; }
      if (A.js[k]) A.js[k][bj] = js;
    }
  }
  if (lane != 0 || sl != 0) return;
  const float px = pa[0], py = pb[0], pz = 0.5f * (pa[1] + pb[2]);   // models/margipose_model.py:254-261
  if (A.coords) { A.coords[bj * 3 + 0] = px; A.coords[bj * 3 + 1] = py; A.coords[bj * 3 + 2] = pz; }
  if (A.loss && A.target) {
    const float dx = px - tx, dy = py - ty, dz = pz - tz;
    float l;
    if (is3d) l = pj[0] + pj[1] + pj[2] + sqrtf(dx * dx + dy * dy + dz * dz);
    else l = pj[0] + sqrtf(dx * dx + dy * dy);
    A.loss[bj] = A.accumulate ? A.loss[bj] + l : l;
  }
}

template <int NV, bool PROJECT, int WPP>
__global__ void __launch_bounds__(WPP > 1 ? 32 * WPP : 256, WPP > 1 ? (WPP >= 16 ? 1 : 16 / WPP) : 2)
tail_bwd_wpj_kernel(const BwdArgs A, const int BJ, const int wshift) {
  pdl_trigger();
  pdl_wait();
  __shared__ float4 red[WPP > 1 ? 32 : 1];
  int phase = 0;
  const int W = A.g.W, HW = A.g.HW;
  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int sl = WPP > 1 ? warp : 0;             // this warp's slice of the plane (WPP > 1: one (sample, joint) per block)
  const int bj = WPP > 1 ? (int)blockIdx.x : blockIdx.x * (blockDim.x >> 5) + warp;
  if (bj >= BJ) return;
  const int w0 = (lane * 4) & (W - 1);
  const int rpi = 128 >> wshift;
  const int h0 = ((sl * NV * 32 + lane) * 4) >> wshift;
  const float col_mult = (float)W * (1.0f / 128.0f), row_mult = 4.0f / (float)W;
  float cw[4];
#pragma unroll
  for (int j = 0; j < 4; ++j) cw[j] = centre(w0 + j, A.g.cw_step, A.g.cw_first);
  const int b = bj / A.J;
  const bool is3d = A.valid_depth ? (A.valid_depth[b] != 0) : true;
  const size_t off = (size_t)bj * HW;
  float tx = 0.f, ty = 0.f, tz = 0.f, w = 0.f, ddx = 0.f, ddy = 0.f, ddz = 0.f;
  if (A.target) {
    tx = A.target[bj * 3 + 0]; ty = A.target[bj * 3 + 1]; tz = A.target[bj * 3 + 2];
    w = A.w[bj];
    const float dx = A.coords[bj * 3 + 0] - tx, dy = A.coords[bj * 3 + 1] - ty;
    const float dz = is3d ? A.coords[bj * 3 + 2] - tz : 0.f;
    const float inv = w / sqrtf(dx * dx + dy * dy + dz * dz);   // infinite at 0, like dsntnn.py:149-150
    ddx = dx * inv; ddy = dy * inv; ddz = dz * inv;
  }
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (!A.out[k]) continue;
    float4 p[NV], d[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      p[i] = __ldg(reinterpret_cast<const float4*>(A.prob[k] + off + ((sl * NV + i) * 32 + lane) * 4));
      d[i] = A.gup[k] ? __ldg(reinterpret_cast<const float4*>(A.gup[k] + off + ((sl * NV + i) * 32 + lane) * 4))
                      : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float wjs, cc, cr, mc, mr;
    bool want_js;
    if (A.coef[k]) {
      wjs = A.coef[k][bj * 3 + 0]; cc = A.coef[k][bj * 3 + 1]; cr = A.coef[k][bj * 3 + 2];
      mc = A.mu[k] ? A.mu[k][bj * 2 + 0] : 0.f;
      mr = A.mu[k] ? A.mu[k][bj * 2 + 1] : 0.f;
      want_js = A.mu[k] != nullptr;
    } else if (A.target) {
      mc = (k == 1) ? tz : tx;
      mr = (k == 2) ? tz : ty;
      wjs = (A.pixelwise && (k == 0 || is3d)) ? w : 0.f;
      cc = (k == 0) ? ddx : (k == 1 ? 0.5f * ddz : 0.f);
      cr = (k == 0) ? ddy : (k == 2 ? 0.5f * ddz : 0.f);
      want_js = wjs != 0.f;
    } else {
      wjs = cc = cr = mc = mr = 0.f;
      want_js = false;
    }
    float ccw[4], qc[4] = {0.f, 0.f, 0.f, 0.f}, qr[NV];
    float ginv = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) ccw[j] = cc * cw[j];
    if (want_js) {
      float sc = 0.f, sr = 0.f;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float dd = cw[j] - mc;
        qc[j] = fast_ex2(__fmul_rn(__fmul_rn(dd, dd), A.g.kw) * L2E);
        sc += qc[j];
      }
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        const float dd = centre(h0 + i * rpi, A.g.ch_step, A.g.ch_first) - mr;
        qr[i] = fast_ex2(__fmul_rn(__fmul_rn(dd, dd), A.g.kh) * L2E);
        sr += qr[i];
      }
      ginv = 1.0f / (warp_sum(sc) * col_mult * (wpj_sum<WPP>(sr, red, phase, warp, lane) * row_mult) + KL_EPS);
    }
    const float kjs = 0.5f * wjs * LN2, hjs = 0.5f * wjs;
    const float TINY = 2e-16f;
    float part = 0.f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float lin_r = cr * centre(h0 + i * rpi, A.g.ch_step, A.g.ch_first);
      const float pv[4] = {p[i].x, p[i].y, p[i].z, p[i].w};
      float dv[4] = {d[i].x, d[i].y, d[i].z, d[i].w};
      if (want_js) {
        const float er = qr[i] * ginv;
        const float lin_js = lin_r + kjs;
        const bool tiny = fminf(fminf(pv[0], pv[1]), fminf(pv[2], pv[3])) < TINY;
        if (!__any_sync(0xffffffffu, tiny)) {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float sq = fmaf(qc[j], er, pv[j]);
            dv[j] += ccw[j] + lin_js;
            dv[j] = fmaf(kjs, fast_lg2(pv[j]) - fast_lg2(sq), dv[j]);
          }
        } else {
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float mm = 0.5f * fmaf(qc[j], er, pv[j]);
            dv[j] += ccw[j] + lin_r;
            dv[j] = fmaf(kjs, fast_lg2(pv[j] + KL_EPS) - fast_lg2(mm + KL_EPS), dv[j]);
            dv[j] = fmaf(hjs, __fdividef(pv[j], pv[j] + KL_EPS) - __fdividef(mm, mm + KL_EPS), dv[j]);
          }
        }
      } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) dv[j] += ccw[j] + lin_r;
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) part = fmaf(pv[j], dv[j], part);
      d[i] = make_float4(dv[0], dv[1], dv[2], dv[3]);
    }
    if (PROJECT) {
      part = wpj_sum<WPP>(part, red, phase, warp, lane);
#pragma unroll
      for (int i = 0; i < NV; ++i) {
        d[i].x = p[i].x * (d[i].x - part); d[i].y = p[i].y * (d[i].y - part);
        d[i].z = p[i].z * (d[i].z - part); d[i].w = p[i].w * (d[i].w - part);
      }
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) *reinterpret_cast<float4*>(A.out[k] + off + ((sl * NV + i) * 32 + lane) * 4) = d[i];
  }
}

// float4 slots per lane of the warp-per-joint kernels (at most max_nv), or 0 when they do not apply
int wpj_slots(int H, int W, int wshift, int max_nv) {
  const int HW = H * W;
  if (wshift < 0 || HW % 128 != 0) return 0;
  const int nv = HW / 128;
  if (nv > max_nv) return 0;
  return (nv == 1 || nv == 2 || nv == 4 || nv == 8 || nv == 16 || nv == 32) ? nv : 0;
}

// Grid of the fast kernels: one CTA per (sample, joint) group by default.  The kernels can also walk several groups
// per CTA (tunable "tail_ctas_per_sm" caps the grid at that many CTAs per SM); measured equal or slower -- the fixed
// per-plane work is not hoistable within 64 registers and a fresh CTA's loads overlap its predecessor's tail.
int persistent_blocks(int groups, int threads) {
  (void)threads;
  if (g_tail_waves <= 0) return groups;
  const int cap = 148 * (int)g_tail_waves;
  return groups < cap ? groups : cap;
}

// log2(W) when the fast kernels apply (W a power of two dividing 128), else -1
int fast_shift(int W) {
  for (int sh = 2; sh <= 7; ++sh)
    if (W == (1 << sh)) return sh;
  return -1;
}

// Chooses warps per plane / pairs per CTA for the warp-sliced kernels; false -> use the block kernels.
bool plan_warps(int HW, int np, int BJ, WarpPlan* P, int* nv) {
  if (np < 1) return false;
  const int vecs = HW / 4;
  P->np = np;
  P->BJ = BJ;
  // Few float4 per lane (NV <= 4) keeps the register count low enough for ~30 resident warps per SM;
  // planes run side by side when np * wpp fits in 16 warps, otherwise one plane at a time.
  for (int cap = (int)g_tail_cap; cap <= 8; cap += 4) {
    for (int wpp = 1; wpp <= 16; ++wpp) {
      const int need = (vecs + wpp * 32 - 1) / (wpp * 32);
      if (need > cap) continue;
      P->wpp = wpp;
      *nv = need <= 1 ? 1 : need <= 2 ? 2 : need <= 4 ? 4 : need <= 6 ? 6 : 8;
      if (np * wpp <= 16) {
        P->seq = 0;
        int groups = 8 / (np * wpp);
        if (groups < 1) groups = 1;
        if (groups > 4) groups = 4;
        P->groups = groups;
      } else {
        P->seq = 1;
        P->groups = 1;
      }
      return true;
    }
  }
  return false;
}

size_t warp_smem(const WarpPlan& P, int H, int W, bool fwd) {
  const int nslots = P.groups * P.np;
  return sizeof(float) * (size_t)(nslots * (fwd ? 2 : 1) * (W + H) + nslots * P.wpp * (fwd ? 5 : 1) +
                                  (fwd ? nslots * 2 : 0) + nslots * P.wpp * 3);   // + normaliser / entropy partials (fast kernels)
}

Geom make_geom(int H, int W, double sigma) {
  Geom g;
  g.H = H; g.W = W; g.HW = H * W;
  g.cw_step = (float)(2.0 / W);
  g.cw_first = (float)(-(W - 1.0) / W);
  g.ch_step = (float)(2.0 / H);
  g.ch_first = (float)(-(H - 1.0) / H);
  const double sw = 2.0 * sigma / W, sh = 2.0 * sigma / H;
  g.kw = (float)(-0.5 * (1.0 / sw) * (1.0 / sw));
  g.kh = (float)(-0.5 * (1.0 / sh) * (1.0 / sh));
  return g;
}

template <bool FROM_LOGITS>
int launch_fwd(const FwdArgs& A, int BJ, bool vec4, cudaStream_t st) {
  const int HW = A.g.HW;
  if (vec4 && g_tail_fast && g_tail_wpj) {
    const int wsh = fast_shift(A.g.W);
    // softmax + expectations only: a plane of up to 4096 elements (64 x 64) still fits one warp's registers (32 float4 per
    // lane, 4 warps per block: 0.93 of the HBM peak at 64 x 64).  With the JS term that variant is latency-bound at 255
    // registers (0.31), so the fused loss keeps the block kernels above 1024 elements.
    const bool with_js = (A.target && A.pixelwise && A.loss) || A.js[0] || A.js[1] || A.js[2];
    const int nv = wpj_slots(A.g.H, A.g.W, wsh, with_js ? 8 : (int)g_tail_wpj_max);
    if (nv) {
      const int wpb = nv > 8 ? 4 : 8;
      const dim3 grid((BJ + wpb - 1) / wpb), block(32 * wpb);
#define MP_FWDJ(NV) mp_launch(tail_fwd_wpj_kernel<NV, FROM_LOGITS, 1>, grid, block, 0, st, A, BJ, wsh)
      if (nv == 1) MP_FWDJ(1); else if (nv == 2) MP_FWDJ(2); else if (nv == 4) MP_FWDJ(4); else if (nv == 8) MP_FWDJ(8);
      else if (nv == 16) MP_FWDJ(16); else MP_FWDJ(32);
#undef MP_FWDJ
      return MP_OK;
    }
    // larger planes: several warps (one block) per plane.  Without a JS term 2 or 4 warps of 32 float4 per lane; with it
    // 8 float4 per lane (the JS pass needs the registers), i.e. 2 ... 16 warps for 2048 ... 16384 elements.
    if (!with_js && wsh >= 0 && g_tail_wpj_max >= 32 && (HW == 2 * 32 * 128 || HW == 4 * 32 * 128)) {
      const dim3 grid(BJ);
      if (HW == 2 * 32 * 128) mp_launch(tail_fwd_wpj_kernel<32, FROM_LOGITS, 2>, grid, dim3(64), 0, st, A, BJ, wsh);
      else mp_launch(tail_fwd_wpj_kernel<32, FROM_LOGITS, 4>, grid, dim3(128), 0, st, A, BJ, wsh);
      return MP_OK;
    }
    if (with_js && wsh >= 0 && g_tail_wpj >= 2 && HW % 1024 == 0) {
      const int wpp = HW / 1024;
      const dim3 grid(BJ), block(32 * wpp);
#define MP_FWDM(WPP) mp_launch(tail_fwd_wpj_kernel<8, FROM_LOGITS, WPP>, grid, block, 0, st, A, BJ, wsh)
      // 16 float4 per lane where the plane is large enough for two warps of them: fewer, longer-lived warps overlap their
      // load / compute / store phases better (128 x 128: 0.81 of the HBM peak against 0.52 with 16 warps of 8 float4,
      // 64 x 64: 0.83 against 0.71; tunable tail_wpj=3 selects the narrow variants)
#define MP_FWDM16(WPP) mp_launch(tail_fwd_wpj_kernel<16, FROM_LOGITS, WPP>, grid, dim3(32 * WPP), 0, st, A, BJ, wsh)
      const bool wide = g_tail_wpj != 3;
      if (wpp == 2) { MP_FWDM(2); return MP_OK; }
      if (wpp == 4) { if (wide) MP_FWDM16(2); else MP_FWDM(4); return MP_OK; }
      if (wpp == 8) { if (wide) MP_FWDM16(4); else MP_FWDM(8); return MP_OK; }
      if (wpp == 16) { if (wide) MP_FWDM16(8); else MP_FWDM(16); return MP_OK; }
#undef MP_FWDM16
#undef MP_FWDM
    }
  }
  if (vec4) {
    WarpPlan P;
    int np = 0, nv = 0;
    for (int k = 0; k < 3; ++k)
      if (A.in[k]) P.pid[np++] = k;
    for (int k = np; k < 3; ++k) P.pid[k] = 0;
    if (plan_warps(HW, np, BJ, &P, &nv)) {
      const int threads = (P.seq ? 1 : P.groups * P.np) * P.wpp * 32;
      const int blocks = (BJ + P.groups - 1) / P.groups;
      const size_t smem = warp_smem(P, A.g.H, A.g.W, true);
      const int wsh = g_tail_fast ? fast_shift(A.g.W) : -1;
      const int pblocks = persistent_blocks(blocks, threads);
#define MP_FWDF(NV) mp_launch(tail_fwd_fast_kernel<NV, FROM_LOGITS>, dim3(pblocks), dim3(threads), smem, st, A, P, wsh)
#define MP_FWDW(NV) mp_launch(tail_fwd_warp_kernel<NV, FROM_LOGITS>, dim3(blocks), dim3(threads), smem, st, A, P)
      if (wsh >= 0) {
        if (nv == 1) MP_FWDF(1); else if (nv == 2) MP_FWDF(2); else if (nv == 4) MP_FWDF(4);
        else if (nv == 6) MP_FWDF(6); else MP_FWDF(8);
      } else {
        if (nv == 1) MP_FWDW(1); else if (nv == 2) MP_FWDW(2); else if (nv == 4) MP_FWDW(4);
        else if (nv == 6) MP_FWDW(6); else MP_FWDW(8);
      }
#undef MP_FWDW
#undef MP_FWDF
      return MP_OK;
    }
  }
#define MP_FWD(VEC, V) mp_launch(tail_fwd_kernel<VEC, V, FROM_LOGITS>, dim3(BJ), dim3(NT), 0, st, A)
  if (vec4) {
    const int v = (HW / 4 + NT - 1) / NT;
    if (v <= 1) MP_FWD(4, 1);
    else if (v <= 2) MP_FWD(4, 2);
    else if (v <= 4) MP_FWD(4, 4);
    else if (v <= 8) MP_FWD(4, 8);
    else if (v <= 16) MP_FWD(4, 16);
    else return MP_ERR_UNSUPPORTED;
  } else {
    const int v = (HW + NT - 1) / NT;
    if (v <= 4) MP_FWD(1, 4);
    else if (v <= 16) MP_FWD(1, 16);
    else if (v <= 64) MP_FWD(1, 64);
    else return MP_ERR_UNSUPPORTED;
  }
#undef MP_FWD
  return MP_OK;
}

template <bool PROJECT>
int launch_bwd(const BwdArgs& A, int BJ, bool vec4, cudaStream_t st) {
  const int HW = A.g.HW;
  if (vec4 && g_tail_fast && g_tail_wpj) {
    const int wsh = fast_shift(A.g.W);
    const int nv = wpj_slots(A.g.H, A.g.W, wsh, 8);   // backward holds p AND the gradient: 1024 elements per warp at most
    if (nv) {
      const dim3 grid((BJ + 7) / 8), block(256);
#define MP_BWDJ(NV) mp_launch(tail_bwd_wpj_kernel<NV, PROJECT, 1>, grid, block, 0, st, A, BJ, wsh)
      if (nv == 1) MP_BWDJ(1); else if (nv == 2) MP_BWDJ(2); else if (nv == 4) MP_BWDJ(4); else MP_BWDJ(8);
#undef MP_BWDJ
      return MP_OK;
    }
    if (wsh >= 0 && g_tail_wpj >= 2 && HW % 1024 == 0) {   // larger planes: 2 ... 16 warps (one block) per plane
      const int wpp = HW / 1024;
      const dim3 grid(BJ), block(32 * wpp);
#define MP_BWDM(WPP) mp_launch(tail_bwd_wpj_kernel<8, PROJECT, WPP>, grid, block, 0, st, A, BJ, wsh)
      if (wpp == 2) { MP_BWDM(2); return MP_OK; }
      if (wpp == 4) { MP_BWDM(4); return MP_OK; }
      if (wpp == 8) { MP_BWDM(8); return MP_OK; }
      if (wpp == 16) { MP_BWDM(16); return MP_OK; }
#undef MP_BWDM
    }
  }
  if (vec4) {
    WarpPlan P;
    int np = 0, nv = 0;
    for (int k = 0; k < 3; ++k)
      if (A.out[k]) P.pid[np++] = k;
    for (int k = np; k < 3; ++k) P.pid[k] = 0;
    if (plan_warps(HW, np, BJ, &P, &nv)) {
      const int threads = (P.seq ? 1 : P.groups * P.np) * P.wpp * 32;
      const int blocks = (BJ + P.groups - 1) / P.groups;
      const size_t smem = warp_smem(P, A.g.H, A.g.W, false);
      const int wsh = g_tail_fast ? fast_shift(A.g.W) : -1;
      size_t smem_f = smem;
      if (wsh >= 0 && PROJECT && nv > 4) {   // gradient stash of the large-plane kernels (see tail_bwd_fast_kernel)
        smem_f = ((smem + 15) & ~(size_t)15) + (size_t)nv * threads * sizeof(float4);
        static bool attr6 = false, attr8 = false;
        if (nv == 6 && !attr6) {
          cudaFuncSetAttribute(tail_bwd_fast_kernel<6, PROJECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
          attr6 = true;
        }
        if (nv == 8 && !attr8) {
          cudaFuncSetAttribute(tail_bwd_fast_kernel<8, PROJECT>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024);
          attr8 = true;
        }
      }
      const int pblocks = persistent_blocks(blocks, threads);
#define MP_BWDF(NV) mp_launch(tail_bwd_fast_kernel<NV, PROJECT>, dim3(pblocks), dim3(threads), smem_f, st, A, P, wsh)
#define MP_BWDW(NV) mp_launch(tail_bwd_warp_kernel<NV, PROJECT>, dim3(blocks), dim3(threads), smem, st, A, P)
      if (wsh >= 0) {
        if (nv == 1) MP_BWDF(1); else if (nv == 2) MP_BWDF(2); else if (nv == 4) MP_BWDF(4);
        else if (nv == 6) MP_BWDF(6); else MP_BWDF(8);
      } else {
        if (nv == 1) MP_BWDW(1); else if (nv == 2) MP_BWDW(2); else if (nv == 4) MP_BWDW(4);
        else if (nv == 6) MP_BWDW(6); else MP_BWDW(8);
      }
#undef MP_BWDW
#undef MP_BWDF
      return MP_OK;
    }
  }
#define MP_BWD(VEC, V) mp_launch(tail_bwd_kernel<VEC, V, PROJECT>, dim3(BJ), dim3(NT), 0, st, A)
  if (vec4) {
    const int v = (HW / 4 + NT - 1) / NT;
    if (v <= 1) MP_BWD(4, 1);
    else if (v <= 2) MP_BWD(4, 2);
    else if (v <= 4) MP_BWD(4, 4);
    else if (v <= 8) MP_BWD(4, 8);
    else if (v <= 16) MP_BWD(4, 16);
    else return MP_ERR_UNSUPPORTED;
  } else {
    const int v = (HW + NT - 1) / NT;
    if (v <= 4) MP_BWD(1, 4);
    else if (v <= 16) MP_BWD(1, 16);
    else if (v <= 64) MP_BWD(1, 64);
    else return MP_ERR_UNSUPPORTED;
  }
#undef MP_BWD
  return MP_OK;
}

bool all_aligned16(const float* const* p, int n) {
  for (int i = 0; i < n; ++i)
    if (p[i] && !mp_aligned16(p[i])) return false;
  return true;
}

}  // namespace

extern "C" {

int mp_tail_fwd(const float* const in[3], int from_logits, float* const prob[3],
                float* const ab[3], float* const js[3], const float* const mu[3],
                const float* target, const int* valid_depth, float* coords, float* loss,
                int accumulate, int pixelwise, double sigma, int B, int J, int H, int W,
                void* stream) {
  MP_CHECK_ARG(in && (in[0] || in[1] || in[2]), "mp_tail_fwd: no input plane");
  MP_CHECK_ARG(B > 0 && J > 0 && H > 0 && W > 0, "mp_tail_fwd: bad shape %dx%dx%dx%d", B, J, H, W);
  MP_CHECK_ARG(H <= MAX_DIM && W <= MAX_DIM, "mp_tail_fwd: H, W must be <= %d", MAX_DIM);
  MP_CHECK_ARG((size_t)H * W <= 16384, "mp_tail_fwd: H*W must be <= 16384");
  MP_CHECK_ARG(!(loss && !target), "mp_tail_fwd: loss output needs targets");
  FwdArgs A;
  for (int k = 0; k < 3; ++k) {
    A.in[k] = in[k];
    A.prob[k] = prob ? prob[k] : nullptr;
    A.ab[k] = ab ? ab[k] : nullptr;
    A.js[k] = js ? js[k] : nullptr;
    A.mu[k] = mu ? mu[k] : nullptr;
    MP_CHECK_ARG(!(A.js[k] && !A.mu[k] && !target), "mp_tail_fwd: js output needs a target");
  }
  A.target = target; A.valid_depth = valid_depth; A.coords = coords; A.loss = loss;
  A.J = J; A.accumulate = accumulate; A.pixelwise = pixelwise;
  A.g = make_geom(H, W, sigma);
  const bool vec4 = (W % 4 == 0) && all_aligned16(A.in, 3) &&
                    all_aligned16((const float* const*)A.prob, 3);
  int rc = from_logits ? launch_fwd<true>(A, B * J, vec4, (cudaStream_t)stream)
                       : launch_fwd<false>(A, B * J, vec4, (cudaStream_t)stream);
  if (rc != MP_OK) { mp_set_error("mp_tail_fwd: unsupported plane size %dx%d", H, W); return rc; }
  MP_CHECK_LAUNCH("mp_tail_fwd");
  return MP_OK;
}

int mp_tail_bwd(const float* const prob[3], const float* const gup[3], float* const out[3],
                const float* target, const float* coords, const float* w,
                const int* valid_depth, const float* const mu[3], const float* const coef[3],
                int project, int pixelwise, double sigma, int B, int J, int H, int W,
                void* stream) {
  MP_CHECK_ARG(prob && out, "mp_tail_bwd: null plane table");
  MP_CHECK_ARG(B > 0 && J > 0 && H > 0 && W > 0, "mp_tail_bwd: bad shape");
  MP_CHECK_ARG(H <= MAX_DIM && W <= MAX_DIM && (size_t)H * W <= 16384, "mp_tail_bwd: plane too large");
  MP_CHECK_ARG(!(target && (!coords || !w)), "mp_tail_bwd: fused mode needs coords and w");
  BwdArgs A;
  for (int k = 0; k < 3; ++k) {
    A.prob[k] = prob[k];
    A.gup[k] = gup ? gup[k] : nullptr;
    A.out[k] = out[k];
    A.mu[k] = mu ? mu[k] : nullptr;
    A.coef[k] = coef ? coef[k] : nullptr;
    MP_CHECK_ARG(!(A.out[k] && !A.prob[k]), "mp_tail_bwd: output plane %d without probabilities", k);
  }
  A.target = target; A.coords = coords; A.w = w; A.valid_depth = valid_depth;
  A.J = J; A.pixelwise = pixelwise;
  A.g = make_geom(H, W, sigma);
  const bool vec4 = (W % 4 == 0) && all_aligned16(A.prob, 3) && all_aligned16(A.gup, 3) &&
                    all_aligned16((const float* const*)A.out, 3);
  int rc = project ? launch_bwd<true>(A, B * J, vec4, (cudaStream_t)stream)
                   : launch_bwd<false>(A, B * J, vec4, (cudaStream_t)stream);
  if (rc != MP_OK) { mp_set_error("mp_tail_bwd: unsupported plane size %dx%d", H, W); return rc; }
  MP_CHECK_LAUNCH("mp_tail_bwd");
  return MP_OK;
}

int mp_euclid_fwd(const float* actual, const float* target, int n, int d, float* out, void* stream) {
  MP_CHECK_ARG(actual && target && out && n >= 0 && d > 0, "mp_euclid_fwd: bad arguments");
  if (n == 0) return MP_OK;
  euclid_fwd_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(actual, target, n, d, out);
  MP_CHECK_LAUNCH("mp_euclid_fwd");
  return MP_OK;
}

int mp_euclid_bwd(const float* grad_out, const float* actual, const float* target,
                  const float* dist, int n, int d, float* grad_actual, void* stream) {
  MP_CHECK_ARG(grad_out && actual && target && dist && grad_actual && n >= 0 && d > 0,
               "mp_euclid_bwd: bad arguments");
  if (n == 0) return MP_OK;
  euclid_bwd_kernel<<<(n + 127) / 128, 128, 0, (cudaStream_t)stream>>>(grad_out, actual, target, dist,
                                                                      n, d, grad_actual);
  MP_CHECK_LAUNCH("mp_euclid_bwd");
  return MP_OK;
}

int mp_make_gauss(const float* mu, float* out, int normalize, double sigma, int BJ, int H, int W,
                  void* stream) {
  MP_CHECK_ARG(mu && out && BJ > 0 && H > 0 && W > 0 && H <= MAX_DIM && W <= MAX_DIM,
               "mp_make_gauss: bad arguments");
  make_gauss_kernel<<<BJ, NT, 0, (cudaStream_t)stream>>>(mu, out, make_geom(H, W, sigma), normalize);
  MP_CHECK_LAUNCH("mp_make_gauss");
  return MP_OK;
}

int mp_masked_mean_fwd(const float* losses, const float* mask, int n, float* out2, void* stream) {
  MP_CHECK_ARG(losses && out2 && n >= 0, "mp_masked_mean_fwd: bad arguments");
  masked_mean_kernel<<<1, NT, 0, (cudaStream_t)stream>>>(losses, mask, n, out2);
  MP_CHECK_LAUNCH("mp_masked_mean_fwd");
  return MP_OK;
}

int mp_masked_mean_bwd(const float* grad_out, const float* mask, const float* mean_den, int n,
                       float* grad_losses, void* stream) {
  MP_CHECK_ARG(grad_out && mean_den && grad_losses && n >= 0, "mp_masked_mean_bwd: bad arguments");
  if (n == 0) return MP_OK;
  masked_mean_bwd_kernel<<<(n + 255) / 256, 256, 0, (cudaStream_t)stream>>>(grad_out, mask, mean_den,
                                                                           n, grad_losses);
  MP_CHECK_LAUNCH("mp_masked_mean_bwd");
  return MP_OK;
}

}  // extern "C"
