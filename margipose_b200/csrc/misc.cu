// Small HBM-bound kernels around the tensor-core path (sm_100a): stem max-pool, the HeatmapColumn
// axis permutation, the HeatmapCombiner 1x1 conv on fp32 NCHW probabilities, the stem im2col
// gather, gradient fan-in sums, fp32->bf16 weight packing and the flat SGD step.
//
// Reference op sites (/root/reference/src/margipose/): models/margipose_model.py:133 (maxpool),
// :86-99 (axis permutation), :142-150,195 (combiner + stage input update), :130 (stem conv1,
// via torchvision), bin/train_3d.py:338-340 (SGD with momentum).
#include "common.cuh"
#include "../../include/margipose_b200.h"

namespace {

// Deterministic mode of the combiner backward: blocks add their weight-gradient partial sums in block order.
// A block waits for its turn; every earlier block has been scheduled before it, so the wait cannot deadlock.
__device__ unsigned int g_combiner_turn = 0;
__device__ __forceinline__ void ordered_begin(int det) {
  if (!det) return;
  if (threadIdx.x == 0) {
    const unsigned me = blockIdx.y * gridDim.x + blockIdx.x;
    while (atomicAdd(&g_combiner_turn, 0u) != me) __nanosleep(64);
  }
  __syncthreads();
}
__device__ __forceinline__ void ordered_end(int det) {
  if (!det) return;
  __threadfence();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned me = blockIdx.y * gridDim.x + blockIdx.x;
    atomicExch(&g_combiner_turn, me + 1 == gridDim.x * gridDim.y ? 0u : me + 1);
  }
}

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

// ------------------------------------------------------------------ maxpool 3x3 / stride 2 / pad 1
__global__ void maxpool_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y,
                                   uint8_t* __restrict__ idx, int N, int H, int W, int C, long long lo_delta) {
  pdl_trigger();
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2, G = C / 8;
  const long long total = (long long)N * Ho * Wo * G;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  // (32-bit index arithmetic: the launcher checks total < 2^31; 64-bit divisions by run-time values cost ~100 instructions each)
  const unsigned iu = (unsigned)i;
  const int g = (int)(iu % (unsigned)G);
  unsigned t = iu / (unsigned)G;
  const int wo = (int)(t % (unsigned)Wo); t /= (unsigned)Wo;
  const int ho = (int)(t % (unsigned)Ho);
  const int n = (int)(t / (unsigned)Ho);
  float best[8];
  int arg[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) { best[j] = -INFINITY; arg[j] = 0; }
  if (lo_delta == 0) {
    // all nine window loads are issued before the first comparison (a guarded load followed by its use, nine times over,
    // is a chain of nine dependent memory round trips)
    uint4 raw[9];
    bool ok[9];
#pragma unroll
    for (int r = 0; r < 3; ++r) {
#pragma unroll
      for (int s = 0; s < 3; ++s) {
        const int h = 2 * ho + r - 1, w = 2 * wo + s - 1;
        ok[r * 3 + s] = h >= 0 && h < H && w >= 0 && w < W;
        raw[r * 3 + s] = ok[r * 3 + s] ? __ldg(reinterpret_cast<const uint4*>(x + (((long long)n * H + h) * W + w) * C + g * 8))
                                       : make_uint4(0u, 0u, 0u, 0u);
      }
    }
#pragma unroll
    for (int t9 = 0; t9 < 9; ++t9) {
      if (!ok[t9]) continue;
      float v[8];
      unpack8(raw[t9], v);
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (v[j] > best[j] || v[j] != v[j]) { best[j] = v[j]; arg[j] = t9; }
    }
  } else {
    for (int r = 0; r < 3; ++r) {
      const int h = 2 * ho + r - 1;
      if (h < 0 || h >= H) continue;
      for (int s = 0; s < 3; ++s) {
        const int w = 2 * wo + s - 1;
        if (w < 0 || w >= W) continue;
        float v[8];
        mp_ld8(x + (((long long)n * H + h) * W + w) * C + g * 8, lo_delta, v);
#pragma unroll
        for (int j = 0; j < 8; ++j)
          if (v[j] > best[j] || v[j] != v[j]) { best[j] = v[j]; arg[j] = r * 3 + s; }
      }
    }
  }
  const long long o = (((long long)n * Ho + ho) * Wo + wo) * C + g * 8;
  mp_st8(y + o, lo_delta, best);
  uint2 a;
  a.x = arg[0] | (arg[1] << 8) | (arg[2] << 16) | (arg[3] << 24);
  a.y = arg[4] | (arg[5] << 8) | (arg[6] << 16) | (arg[7] << 24);
  *reinterpret_cast<uint2*>(idx + o) = a;
}

__global__ void maxpool_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const uint8_t* __restrict__ idx,
                                   __nv_bfloat16* __restrict__ dx, int N, int H, int W, int C) {
  pdl_trigger();
  pdl_wait();
  const int Ho = H / 2, Wo = W / 2, G = C / 8;
  const long long total = (long long)N * H * W * G;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const unsigned iu = (unsigned)i;   // (32-bit index arithmetic, see maxpool_fwd_kernel)
  const int g = (int)(iu % (unsigned)G);
  unsigned t = iu / (unsigned)G;
  const int w = (int)(t % (unsigned)W); t /= (unsigned)W;
  const int h = (int)(t % (unsigned)H);
  const int n = (int)(t / (unsigned)H);
  float acc[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) acc[j] = 0.f;
  // output windows (ho, wo) that contain (h, w): 2*ho + r - 1 == h, i.e. r = 1 for even h, r in {0, 2} for odd h (same for
  // columns): at most four candidates, all loaded before the first one is used
  const int rr[2] = {(h & 1) ? 0 : 1, (h & 1) ? 2 : -1};
  const int ss[2] = {(w & 1) ? 0 : 1, (w & 1) ? 2 : -1};
  uint2 ia[4];
  uint4 va[4];
  int tapv[4];
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    const int r = rr[q >> 1], sx = ss[q & 1];
    const int hh = h + 1 - r, ww = w + 1 - sx;
    const bool in = r >= 0 && sx >= 0 && hh >= 0 && ww >= 0 && (hh >> 1) < Ho && (ww >> 1) < Wo;
    tapv[q] = in ? r * 3 + sx : -1;
    const long long o = in ? (((long long)n * Ho + (hh >> 1)) * Wo + (ww >> 1)) * C + g * 8 : 0;
    ia[q] = in ? __ldg(reinterpret_cast<const uint2*>(idx + o)) : make_uint2(0xffffffffu, 0xffffffffu);
    va[q] = in ? __ldg(reinterpret_cast<const uint4*>(dy + o)) : make_uint4(0u, 0u, 0u, 0u);
  }
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    float v[8];
    unpack8(va[q], v);
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int aj = (int)(((j < 4 ? ia[q].x : ia[q].y) >> ((j & 3) * 8)) & 0xff);
      if (aj == tapv[q]) acc[j] += v[j];
    }
  }
  *reinterpret_cast<uint4*>(dx + (((long long)n * H + h) * W + w) * C + g * 8) = pack8(acc);
}

// ------------------------------------------------------------------------------ axis permutation
// out[n, h, c', g*S + w] = in[n, h, w, g*S + c']  (mode 1)
// out[n, c', w, g*S + h] = in[n, h, w, g*S + c']  (mode 2)
__global__ void axis_permute_kernel(const __nv_bfloat16* __restrict__ in, __nv_bfloat16* __restrict__ out,
                                    int mode, int N, int S, int C, int Cp) {
  pdl_trigger();
  pdl_wait();
  const long long total = (long long)N * S * S * Cp;
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const unsigned iu = (unsigned)i;   // (32-bit index arithmetic, see maxpool_fwd_kernel)
  const int ch = (int)(iu % (unsigned)Cp);
  unsigned t = iu / (unsigned)Cp;
  const int col = (int)(t % (unsigned)S); t /= (unsigned)S;
  const int row = (int)(t % (unsigned)S);
  const int n = (int)(t / (unsigned)S);
  __nv_bfloat16 v = __float2bfloat16(0.f);
  if (ch < C) {
    const int g = ch / S, k = ch - g * S;
    // out(row, col, g*S + k):  mode 1: h = row, c' = col, w = k ;  mode 2: c' = row, w = col, h = k
    int h, w, c;
    if (mode == 1) { h = row; w = k; c = col; }
    else { h = k; w = col; c = row; }
    v = in[(((long long)n * S + h) * S + w) * Cp + g * S + c];
  }
  out[i] = v;
}

// ------------------------------------------------------------------------------------ combiner
// One block = 32 pixels of one image: probabilities staged in shared memory, each thread owns
// 8 consecutive output channels of 2 pixels... kept simple: thread = (pixel, 8 channels).
// ---- register-tiled combiner kernels (C <= 128): 64 pixels per tile, 128-bit shared-memory loads, ~10 FMA per load
constexpr int CT_PIX = 64;

// out[pix, c] = inp[pix, c] + sum_k w[c, k] * p_k[pix]; thread = 8 channels x 4 pixels.
__global__ void __launch_bounds__(256) combiner_fwd_tiled_kernel(const float* __restrict__ p0, const float* __restrict__ p1,
                                                               const float* __restrict__ p2, const float* __restrict__ w,
                                                               const __nv_bfloat16* __restrict__ inp,
                                                               __nv_bfloat16* __restrict__ out, int J, int HW, int C,
                                                               long long lo_delta) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float smt[];
  const int K = 3 * J, CS = C + 4, G = C / 8;
  float* swT = smt;              // [K][C + 4]: k-major so a thread's 8 channels are two 128-bit loads
  float* sp = smt + K * CS;      // [K][64]
  const int n = blockIdx.y, pix0 = blockIdx.x * CT_PIX;
  for (int i = threadIdx.x; i < C * K; i += blockDim.x) {
    const int c = i / K, k = i - c * K;
    swT[k * CS + c] = w[i];
  }
  for (int i = threadIdx.x; i < K * CT_PIX; i += blockDim.x) {
    const int k = i / CT_PIX, px = i - k * CT_PIX;
    const float* src = k < J ? p0 : (k < 2 * J ? p1 : p2);
    sp[i] = (pix0 + px < HW) ? src[((long long)n * J + (k % J)) * HW + pix0 + px] : 0.f;
  }
  __syncthreads();
  const int g = threadIdx.x % G, q = threadIdx.x / G;   // 16 pixel quads x G channel groups
  if (q >= CT_PIX / 4) return;
  float acc[4][8];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;
  for (int k = 0; k < K; ++k) {
    const float4 wa = *reinterpret_cast<const float4*>(swT + k * CS + g * 8);
    const float4 wb = *reinterpret_cast<const float4*>(swT + k * CS + g * 8 + 4);
    const float4 pv = *reinterpret_cast<const float4*>(sp + k * CT_PIX + q * 4);
    const float wv[8] = {wa.x, wa.y, wa.z, wa.w, wb.x, wb.y, wb.z, wb.w};
    const float pp[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(wv[j], pp[i], acc[i][j]);
  }
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const int px = pix0 + q * 4 + i;
    if (px >= HW) break;
    const long long o = ((long long)n * HW + px) * C + g * 8;
    float base[8];
    mp_ld8(inp + o, lo_delta, base);
    // fp32 weights, fp32 probabilities, fp32 accumulate; the new stage input is rounded to bf16 once
#pragma unroll
    for (int j = 0; j < 8; ++j) base[j] += acc[i][j];
    mp_st8(out + o, lo_delta, base);
  }
}

// dp_k[n, j, pix] = sum_c w[c, k] * dout[pix, c]  (thread = 4 pixels x 4 k);
// dw[c, k] += sum_pix dout[pix, c] * p_k[pix]      (thread = 8 channels x 4 k, kept in registers over the block's tiles)
constexpr int CT_TILES = 2;
__global__ void __launch_bounds__(256) combiner_bwd_tiled_kernel(const __nv_bfloat16* __restrict__ dout,
                                                               const float* __restrict__ p0, const float* __restrict__ p1,
                                                               const float* __restrict__ p2, const float* __restrict__ w,
                                                               float* __restrict__ dp0, float* __restrict__ dp1,
                                                               float* __restrict__ dp2, float* __restrict__ dw,
                                                               int accumulate, int J, int HW, int C, long long lo_delta,
                                                               int det) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ __align__(16) float smt[];
  const int K = 3 * J, KP = (K + 3) & ~3, CS = C + 4, G = C / 8, KG = KP / 4;
  float* sdT = smt;                    // [64][C + 4] upstream gradient tile, fp32
  float* sw = sdT + CT_PIX * CS;       // [C][KP]
  float* spT = sw + C * KP;            // [64][KP] probabilities, pixel-major
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < C * KP; i += blockDim.x) {
    const int c = i / KP, k = i - c * KP;
    sw[i] = k < K ? w[c * K + k] : 0.f;
  }
  const int kq = threadIdx.x % 16, q = threadIdx.x / 16;      // dp role: k group (4 k) x pixel quad
  const int cg = threadIdx.x % G, kg = threadIdx.x / G;       // dw role: channel group (8 c) x k group (4 k)
  const bool dw_active = kg < KG;
  float wacc[8][4];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) wacc[i][j] = 0.f;
  for (int tile = 0; tile < CT_TILES; ++tile) {
    const int pix0 = (blockIdx.x * CT_TILES + tile) * CT_PIX;
    if (pix0 >= HW) break;
    __syncthreads();
    for (int i = threadIdx.x; i < CT_PIX * G; i += blockDim.x) {
      const int px = i / G, g = i - px * G;
      float v[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) v[j] = 0.f;
      if (pix0 + px < HW) mp_ld8(dout + ((long long)n * HW + pix0 + px) * C + g * 8, lo_delta, v);
      *reinterpret_cast<float4*>(sdT + px * CS + g * 8) = make_float4(v[0], v[1], v[2], v[3]);
      *reinterpret_cast<float4*>(sdT + px * CS + g * 8 + 4) = make_float4(v[4], v[5], v[6], v[7]);
    }
    for (int i = threadIdx.x; i < KP * CT_PIX; i += blockDim.x) {
      const int k = i / CT_PIX, px = i - k * CT_PIX;
      float v = 0.f;
      if (k < K && pix0 + px < HW) {
        const float* src = k < J ? p0 : (k < 2 * J ? p1 : p2);
        v = src[((long long)n * J + (k % J)) * HW + pix0 + px];
      }
      spT[px * KP + k] = v;
    }
    __syncthreads();
    if (kq * 4 < K) {   // ---- dp: 4 pixels x 4 k per thread
      float acc[4][4];
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = 0.f;
      for (int c = 0; c < C; c += 4) {
        float d[4][4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float4 t = *reinterpret_cast<const float4*>(sdT + (q * 4 + i) * CS + c);
          d[i][0] = t.x; d[i][1] = t.y; d[i][2] = t.z; d[i][3] = t.w;
        }
#pragma unroll
        for (int cc = 0; cc < 4; ++cc) {
          const float4 t = *reinterpret_cast<const float4*>(sw + (c + cc) * KP + kq * 4);
          const float wv[4] = {t.x, t.y, t.z, t.w};
#pragma unroll
          for (int i = 0; i < 4; ++i)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(wv[j], d[i][cc], acc[i][j]);
        }
      }
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = kq * 4 + j;
        if (k >= K) break;
        float* dst = (k < J ? dp0 : (k < 2 * J ? dp1 : dp2)) + ((long long)n * J + (k % J)) * HW + pix0 + q * 4;
#pragma unroll
        for (int i = 0; i < 4; ++i)
          if (pix0 + q * 4 + i < HW) dst[i] = accumulate ? dst[i] + acc[i][j] : acc[i][j];
      }
    }
    if (dw_active) {   // ---- dw: 8 channels x 4 k per thread over the tile's pixels
      for (int px = 0; px < CT_PIX; ++px) {
        const float4 da = *reinterpret_cast<const float4*>(sdT + px * CS + cg * 8);
        const float4 db = *reinterpret_cast<const float4*>(sdT + px * CS + cg * 8 + 4);
        const float4 pv = *reinterpret_cast<const float4*>(spT + px * KP + kg * 4);
        const float dv[8] = {da.x, da.y, da.z, da.w, db.x, db.y, db.z, db.w};
        const float pp[4] = {pv.x, pv.y, pv.z, pv.w};
#pragma unroll
        for (int i = 0; i < 8; ++i)
#pragma unroll
          for (int j = 0; j < 4; ++j) wacc[i][j] = fmaf(dv[i], pp[j], wacc[i][j]);
      }
    }
  }
  ordered_begin(det);
  if (dw_active) {
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int k = kg * 4 + j;
        if (k < K) atomicAdd(dw + (cg * 8 + i) * K + k, wacc[i][j]);
      }
  }
  ordered_end(det);
}

constexpr int CMB_PIX = 32;
__global__ void __launch_bounds__(256) combiner_fwd_kernel(const float* __restrict__ p0, const float* __restrict__ p1,
                                                         const float* __restrict__ p2, const float* __restrict__ w,
                                                         const __nv_bfloat16* __restrict__ inp,
                                                         __nv_bfloat16* __restrict__ out, int J, int HW, int C,
                                                         long long lo_delta) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  const int K = 3 * J;
  float* sp = sm;                 // [K][CMB_PIX]
  float* sw = sm + K * CMB_PIX;   // [C][K]
  const int n = blockIdx.y;
  const int pix0 = blockIdx.x * CMB_PIX;
  for (int i = threadIdx.x; i < K * CMB_PIX; i += blockDim.x) {
    const int k = i / CMB_PIX, px = i - k * CMB_PIX;
    const float* src = k < J ? p0 : (k < 2 * J ? p1 : p2);
    const int j = k % J;
    sp[i] = (pix0 + px < HW) ? src[((long long)n * J + j) * HW + pix0 + px] : 0.f;
  }
  for (int i = threadIdx.x; i < C * K; i += blockDim.x) sw[i] = w[i];
  __syncthreads();
  const int G = C / 8;
  for (int i = threadIdx.x; i < CMB_PIX * G; i += blockDim.x) {
    const int px = i / G, g = i - px * G;
    if (pix0 + px >= HW) continue;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int k = 0; k < K; ++k) {
      const float pv = sp[k * CMB_PIX + px];
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] = fmaf(sw[(g * 8 + j) * K + k], pv, acc[j]);
    }
    const long long o = ((long long)n * HW + pix0 + px) * C + g * 8;
    float base[8];
    mp_ld8(inp + o, lo_delta, base);
    // fp32 weights, fp32 probabilities, fp32 accumulate; the new stage input is rounded to bf16 once
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = base[j] + acc[j];
    mp_st8(out + o, lo_delta, acc);
  }
}

// dp_k[n, j, pix] = sum_c w[c, k*J + j] * dout[pix, c];  dw[c, kk] += sum_pix dout[pix, c] * p[kk, pix]
// One block walks CMB_TILES tiles of 32 pixels of one image; the weight-gradient partial sums stay in
// registers across the tiles, so the block issues one atomic per weight at the end.
constexpr int CMB_TILES = 2;   // few tiles per block: 4 blocks (32 warps) per SM hide the shared-memory latency of the FMA loops
constexpr int CMB_WPT = 26;   // weights per thread: ceil(128 * 51 / 256)
__global__ void __launch_bounds__(256) combiner_bwd_kernel(const __nv_bfloat16* __restrict__ dout,
                                                         const float* __restrict__ p0, const float* __restrict__ p1,
                                                         const float* __restrict__ p2, const float* __restrict__ w,
                                                         float* __restrict__ dp0, float* __restrict__ dp1,
                                                         float* __restrict__ dp2, float* __restrict__ dw,
                                                         int accumulate, int J, int HW, int C, long long lo_delta,
                                                         int det) {
  pdl_trigger();
  pdl_wait();
  extern __shared__ float sm[];
  const int K = 3 * J;
  float* sd = sm;                    // [CMB_PIX][C + 1]
  float* sw = sd + CMB_PIX * (C + 1);   // [C][K]
  float* sp = sw + C * K;            // [K][CMB_PIX + 1] (padded: the dW loop walks k across lanes)
  const int n = blockIdx.y;
  for (int i = threadIdx.x; i < C * K; i += blockDim.x) sw[i] = w[i];
  float wacc[CMB_WPT];
#pragma unroll
  for (int q = 0; q < CMB_WPT; ++q) wacc[q] = 0.f;
  for (int tile = 0; tile < CMB_TILES; ++tile) {
    const int pix0 = (blockIdx.x * CMB_TILES + tile) * CMB_PIX;
    if (pix0 >= HW) break;
    __syncthreads();
    for (int i = threadIdx.x; i < CMB_PIX * C; i += blockDim.x) {
      const int px = i / C, c = i - px * C;
      float dv = 0.f;
      if (pix0 + px < HW) {
        const long long o = ((long long)n * HW + pix0 + px) * C + c;
        dv = __bfloat162float(dout[o]);
        if (lo_delta) dv += __bfloat162float(dout[o + lo_delta]);
      }
      sd[px * (C + 1) + c] = dv;
    }
    for (int i = threadIdx.x; i < K * CMB_PIX; i += blockDim.x) {
      const int k = i / CMB_PIX, px = i - k * CMB_PIX;
      const float* src = k < J ? p0 : (k < 2 * J ? p1 : p2);
      sp[k * (CMB_PIX + 1) + px] = (pix0 + px < HW) ? src[((long long)n * J + (k % J)) * HW + pix0 + px] : 0.f;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < K * CMB_PIX; i += blockDim.x) {
      const int k = i / CMB_PIX, px = i - k * CMB_PIX;
      if (pix0 + px >= HW) continue;
      float acc = 0.f;
      for (int c = 0; c < C; ++c) acc = fmaf(sw[c * K + k], sd[px * (C + 1) + c], acc);
      float* dst = k < J ? dp0 : (k < 2 * J ? dp1 : dp2);
      float* d = dst + ((long long)n * J + (k % J)) * HW + pix0 + px;
      *d = accumulate ? *d + acc : acc;
    }
#pragma unroll
    for (int q = 0; q < CMB_WPT; ++q) {
      const int i = threadIdx.x + q * 256;
      if (i < C * K) {
        const int c = i / K, k = i - c * K;
        float acc = wacc[q];
#pragma unroll 8
        for (int px = 0; px < CMB_PIX; ++px) acc = fmaf(sd[px * (C + 1) + c], sp[k * (CMB_PIX + 1) + px], acc);
        wacc[q] = acc;
      }
    }
  }
  ordered_begin(det);
#pragma unroll
  for (int q = 0; q < CMB_WPT; ++q) {
    const int i = threadIdx.x + q * 256;
    if (i < C * K) atomicAdd(dw + i, wacc[q]);
  }
  ordered_end(det);
}

// -------------------------------------------------------------------------------- stem im2col
// patches[n, ho, wo, (r*7+s)*3 + c] = x[n, c, 2*ho + r - 3, 2*wo + s - 3]; 192 columns per row (147 real).  From the fp32
// NCHW image, or from a uint8 NHWC image (what an image decoder produces) with the input step of the reference fused in:
// to_tensor (/255) and ImageNet normalisation (x - mean) / std (data_specs.py:6-13,38-39); padding taps are zeros of the
// NORMALISED tensor, as the conv's zero padding sees them.
// A block owns STEM_TW consecutive output pixels of one output row, stages the 7 x (2 * STEM_TW + 5) x 3 input window it needs
// in shared memory with coalesced reads (normalised, zero outside the image), and writes the pixels' 192-column patch rows
// -- one contiguous run of STEM_TW * 384 bytes -- with fully coalesced 16-byte stores.  (A thread per (pixel, 8 columns) that
// gathers straight from global memory stores 16 bytes per 384-byte row and lane: 187 us instead of 124 us per step.)
constexpr int STEM_TW = 32, STEM_WW = 2 * STEM_TW + 5;
template <bool U8>
__global__ void __launch_bounds__(192) stem_im2col_tiled_kernel(const void* __restrict__ xin, __nv_bfloat16* __restrict__ out,
                                                                int N, int H, int W, float3 scale, float3 shift,
                                                                long long lo_delta) {
  pdl_trigger();
  pdl_wait();
  __shared__ float win[3][7][STEM_WW + 3];
  const int Ho = H / 2, Wo = W / 2;
  const int tiles_w = (Wo + STEM_TW - 1) / STEM_TW;
  const int tw = blockIdx.x % tiles_w;
  const int ho = (blockIdx.x / tiles_w) % Ho;
  const int n = blockIdx.x / (tiles_w * Ho);
  const int wo0 = tw * STEM_TW, w_first = 2 * wo0 - 3, h_first = 2 * ho - 3;
#pragma unroll 4
  for (int i = threadIdx.x; i < 3 * 7 * STEM_WW; i += blockDim.x) {
    int c, r, col;
    if (U8) {   // NHWC bytes: channel fastest
      c = i % 3; col = (i / 3) % STEM_WW; r = i / (3 * STEM_WW);
    } else {    // NCHW floats: column fastest
      col = i % STEM_WW; r = (i / STEM_WW) % 7; c = i / (7 * STEM_WW);
    }
    const int h = h_first + r, w = w_first + col;
    float val = 0.f;
    if (h >= 0 && h < H && w >= 0 && w < W) {
      if (U8) {
        const float px = (float)__ldg(reinterpret_cast<const uint8_t*>(xin) + (((long long)n * H + h) * W + w) * 3 + c);
        val = fmaf(px, c == 0 ? scale.x : (c == 1 ? scale.y : scale.z), c == 0 ? shift.x : (c == 1 ? shift.y : shift.z));
      } else {
        val = __ldg(reinterpret_cast<const float*>(xin) + (((long long)n * 3 + c) * H + h) * W + w);
      }
    }
    win[c][r][col] = val;
  }
  __syncthreads();
  // 192 threads = 8 pixels x 24 column groups per pass: a thread keeps ITS column group, so the window offsets of its
  // eight columns are computed once; consecutive threads store consecutive 16-byte pieces of the row
  __nv_bfloat16* row = out + (((long long)n * Ho + ho) * Wo + wo0) * 192;
  const int g = threadIdx.x % 24, p0 = threadIdx.x / 24;
  const float* wf = &win[0][0][0];
  int offs[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = g * 8 + j;
    const int tap = k / 3, c = k - tap * 3;
    const int r = tap / 7, sx = tap - r * 7;
    offs[j] = k < 147 ? (c * 7 + r) * (STEM_WW + 3) + sx : -1;
  }
  for (int px = p0; px < STEM_TW; px += 8) {
    if (wo0 + px >= Wo) break;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) v[j] = offs[j] >= 0 ? wf[offs[j] + 2 * px] : 0.f;
    mp_st8(row + (long long)(px * 24 + g) * 8, lo_delta, v);
  }
}

// -------------------------------------------------------------------------------------- add_n
__global__ void add_bf16_kernel(const __nv_bfloat16* a, const __nv_bfloat16* b, const __nv_bfloat16* c,
                                const __nv_bfloat16* d, int n, __nv_bfloat16* out, long long groups,
                                long long lo_delta) {
  pdl_trigger();
  pdl_wait();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= groups) return;
  float acc[8], v[8];
  mp_ld8(a + i * 8, lo_delta, acc);
  const __nv_bfloat16* rest[3] = {b, c, d};
#pragma unroll
  for (int k = 0; k < 3; ++k) {
    if (k + 1 < n) {
      mp_ld8(rest[k] + i * 8, lo_delta, v);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc[j] += v[j];
    }
  }
  mp_st8(out + i * 8, lo_delta, acc);
}

// ------------------------------------------------------------------------------- weight packing
__global__ void pack_weights_kernel(const float* __restrict__ master, __nv_bfloat16* __restrict__ packed,
                                    const mp_pack_entry* __restrict__ table, int n_entries, long long groups) {
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= groups) return;
  const long long e0 = i * 8;   // first work element of this thread's 8
  int lo = 0, hi = n_entries - 1;
  while (lo < hi) {   // first entry with work_end > e0
    const int mid = (lo + hi) >> 1;
    if (table[mid].work_end > e0) hi = mid; else lo = mid + 1;
  }
  const mp_pack_entry E = table[lo];
  const unsigned local = (unsigned)((e0 - E.work_off) >> 3);   // this thread's 8-column group within the entry
  const unsigned cgroups = (unsigned)E.cols_p >> 3;
  int c0, t, r;
  if (!E.transpose) {   // groups ordered (r, t, c-group): source rows are read along c, contiguously
    c0 = (int)(local % cgroups) * 8;
    const unsigned rt = local / cgroups;
    t = (int)(rt % (unsigned)E.taps);
    r = (int)(rt / (unsigned)E.taps);
  } else {              // groups ordered (c-group, t, r): consecutive threads read consecutive r of the source
    r = (int)(local % (unsigned)E.rows_p);
    const unsigned ct = local / (unsigned)E.rows_p;
    t = (int)(ct % (unsigned)E.taps);
    c0 = (int)(ct / (unsigned)E.taps) * 8;
  }
  float v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = c0 + j;
    float val = 0.f;
    if (!E.transpose) {
      if (r < E.A && c < E.B) val = master[E.src_off + ((long long)r * E.taps + t) * E.B + c];
    } else {
      if (c < E.A && r < E.B) val = master[E.src_off + ((long long)c * E.taps + t) * E.B + r];
    }
    v[j] = val;
  }
  mp_st8(packed + E.dst_off + (long long)r * E.dst_row_stride + (long long)t * E.cols_p + c0, E.lo_off, v);
}

// ------------------------------------------------------------------------------------------ SGD
__device__ __forceinline__ float sgd_one(float w, float grad, float& b, float lr, float mom, float damp, float wd,
                                        int nesterov, int first) {
  if (wd != 0.f) grad = fmaf(wd, w, grad);
  if (mom != 0.f) {
    b = first ? grad : fmaf(mom, b, (1.f - damp) * grad);
    grad = nesterov ? fmaf(mom, b, grad) : b;
  }
  return fmaf(-lr, grad, w);
}

// 4 parameters per thread (128-bit accesses), grid-stride; the scalar tail goes to the last few threads.
__global__ void __launch_bounds__(256) sgd_kernel(float* __restrict__ p, const float* __restrict__ g,
                                                  float* __restrict__ buf, long long n, float lr, float mom, float damp,
                                                  float wd, int nesterov, int first, float gscale,
                                                  const float* __restrict__ hyper) {
  if (hyper) {   // hyperparameters live in device memory: schedulers keep working when the step is a replayed CUDA graph
    lr = hyper[0]; mom = hyper[1]; damp = hyper[2]; wd = hyper[3]; gscale = hyper[4];
    first = hyper[5] != 0.f;   // "first step initialises the momentum buffer" follows the host's step count too
  }
  const long long n4 = n >> 2;
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    float4 w = reinterpret_cast<float4*>(p)[i];
    const float4 gr = __ldg(reinterpret_cast<const float4*>(g) + i);
    float4 b = make_float4(0.f, 0.f, 0.f, 0.f);
    if (mom != 0.f && !first) b = reinterpret_cast<float4*>(buf)[i];
    w.x = sgd_one(w.x, gr.x * gscale, b.x, lr, mom, damp, wd, nesterov, first);
    w.y = sgd_one(w.y, gr.y * gscale, b.y, lr, mom, damp, wd, nesterov, first);
    w.z = sgd_one(w.z, gr.z * gscale, b.z, lr, mom, damp, wd, nesterov, first);
    w.w = sgd_one(w.w, gr.w * gscale, b.w, lr, mom, damp, wd, nesterov, first);
    if (mom != 0.f) reinterpret_cast<float4*>(buf)[i] = b;
    reinterpret_cast<float4*>(p)[i] = w;
  }
  const long long t = (n4 << 2) + (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n) {
    float b = (mom != 0.f && !first) ? buf[t] : 0.f;
    p[t] = sgd_one(p[t], g[t] * gscale, b, lr, mom, damp, wd, nesterov, first);
    if (mom != 0.f) buf[t] = b;
  }
}

}  // namespace

extern "C" {

#define MP_CHECK_DELTA(what) \
  MP_CHECK_ARG(lo_delta >= 0 && lo_delta % 8 == 0, what ": lo_delta must be a non-negative multiple of 8")

int mp_maxpool_fwd(const void* x, void* y, uint8_t* idx, int N, int H, int W, int C, int64_t lo_delta, void* stream) {
  MP_CHECK_DELTA("mp_maxpool_fwd");
  MP_CHECK_ARG(x && y && idx && N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0,
               "mp_maxpool_fwd: bad arguments");
  const long long total = (long long)N * (H / 2) * (W / 2) * (C / 8);
  MP_CHECK_ARG(total < (1LL << 31), "mp_maxpool_fwd: tensor too large (32-bit element index)");
  MP_CUDA(mp_launch(maxpool_fwd_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)x, (__nv_bfloat16*)y, idx, N, H, W, C, (long long)lo_delta));
  MP_CHECK_LAUNCH("mp_maxpool_fwd");
  return MP_OK;
}

int mp_maxpool_bwd(const void* dy, const uint8_t* idx, void* dx, int N, int H, int W, int C, void* stream) {
  MP_CHECK_ARG(dy && dx && idx && N > 0 && H % 2 == 0 && W % 2 == 0 && C % 8 == 0, "mp_maxpool_bwd: bad arguments");
  const long long total = (long long)N * H * W * (C / 8);
  MP_CHECK_ARG(total < (1LL << 31), "mp_maxpool_bwd: tensor too large (32-bit element index)");
  MP_CUDA(mp_launch(maxpool_bwd_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)dy, idx, (__nv_bfloat16*)dx, N, H, W, C));
  MP_CHECK_LAUNCH("mp_maxpool_bwd");
  return MP_OK;
}

int mp_axis_permute(const void* in, void* out, int mode, int N, int S, int C, int Cp, void* stream) {
  MP_CHECK_ARG(in && out && in != out && (mode == 1 || mode == 2) && N > 0 && S > 0 && C % S == 0 && Cp >= C,
               "mp_axis_permute: bad arguments (the spatial size must divide the channel count)");
  const long long total = (long long)N * S * S * Cp;
  MP_CHECK_ARG(total < (1LL << 31), "mp_axis_permute: tensor too large (32-bit element index)");
  MP_CUDA(mp_launch(axis_permute_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)in, (__nv_bfloat16*)out, mode, N, S, C, Cp));
  MP_CHECK_LAUNCH("mp_axis_permute");
  return MP_OK;
}

int mp_combiner_fwd(const float* const p[3], const float* w, const void* inp, void* out, int N, int J, int HW,
                    int C, int64_t lo_delta, void* stream) {
  MP_CHECK_DELTA("mp_combiner_fwd");
  MP_CHECK_ARG(p && p[0] && p[1] && p[2] && w && inp && out && N > 0 && J > 0 && HW > 0 && C % 8 == 0,
               "mp_combiner_fwd: bad arguments");
  if (C <= 128 && 3 * J <= 64 && 256 % (C / 8) == 0) {   // register-tiled kernel (the MargiPose shape: C = 128, J = 17)
    const size_t smem_t = (size_t)(3 * J * (C + 4) + 3 * J * CT_PIX) * sizeof(float);
    static bool attr_f = false;
    if (!attr_f) {
      MP_CUDA(cudaFuncSetAttribute(combiner_fwd_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr_f = true;
    }
    dim3 grid_t((HW + CT_PIX - 1) / CT_PIX, N);
    MP_CUDA(mp_launch(combiner_fwd_tiled_kernel, grid_t, dim3(256), smem_t, (cudaStream_t)stream, p[0], p[1], p[2], w,
                      (const __nv_bfloat16*)inp, (__nv_bfloat16*)out, J, HW, C, (long long)lo_delta));
    MP_CHECK_LAUNCH("mp_combiner_fwd");
    return MP_OK;
  }
  const size_t smem = (size_t)(3 * J * CMB_PIX + C * 3 * J) * sizeof(float);
  MP_CHECK_ARG(smem <= 48 * 1024, "mp_combiner_fwd: J*C too large");
  dim3 grid((HW + CMB_PIX - 1) / CMB_PIX, N);
  MP_CUDA(mp_launch(combiner_fwd_kernel, dim3(grid), dim3(256), smem, (cudaStream_t)stream, p[0], p[1], p[2], w, (const __nv_bfloat16*)inp,
                                                                 (__nv_bfloat16*)out, J, HW, C, (long long)lo_delta));
  MP_CHECK_LAUNCH("mp_combiner_fwd");
  return MP_OK;
}

int mp_combiner_bwd(const void* dout, const float* const p[3], const float* w, float* const dp[3], float* dw,
                    int accumulate, int N, int J, int HW, int C, int64_t lo_delta, void* stream) {
  MP_CHECK_DELTA("mp_combiner_bwd");
  MP_CHECK_ARG(dout && p && p[0] && p[1] && p[2] && w && dp && dp[0] && dp[1] && dp[2] && dw && N > 0 && J > 0 &&
                   HW > 0 && C > 0,
               "mp_combiner_bwd: bad arguments");
  if (C % 8 == 0 && C <= 128 && 3 * J <= 64 && 256 % (C / 8) == 0 && (256 / (C / 8)) * 4 >= ((3 * J + 3) & ~3)) {
    const int KP = (3 * J + 3) & ~3;
    const size_t smem_t = (size_t)(CT_PIX * (C + 4) + C * KP + CT_PIX * KP) * sizeof(float);
    static bool attr_b = false;
    if (!attr_b) {
      MP_CUDA(cudaFuncSetAttribute(combiner_bwd_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr_b = true;
    }
    dim3 grid_t((HW + CT_PIX * CT_TILES - 1) / (CT_PIX * CT_TILES), N);
    MP_CUDA(mp_launch(combiner_bwd_tiled_kernel, grid_t, dim3(256), smem_t, (cudaStream_t)stream,
                      (const __nv_bfloat16*)dout, p[0], p[1], p[2], w, dp[0], dp[1], dp[2], dw, accumulate, J, HW, C,
                      (long long)lo_delta, mp_deterministic()));
    MP_CHECK_LAUNCH("mp_combiner_bwd");
    return MP_OK;
  }
  const size_t smem = (size_t)(CMB_PIX * (C + 1) + C * 3 * J + 3 * J * (CMB_PIX + 1)) * sizeof(float);
  MP_CHECK_ARG(smem <= 96 * 1024 && (long long)C * 3 * J <= 256 * CMB_WPT, "mp_combiner_bwd: J*C too large");
  static bool attr_set = false;
  if (!attr_set) {
    MP_CUDA(cudaFuncSetAttribute(combiner_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set = true;
  }
  dim3 grid((HW + CMB_PIX * CMB_TILES - 1) / (CMB_PIX * CMB_TILES), N);
  MP_CUDA(mp_launch(combiner_bwd_kernel, dim3(grid), dim3(256), smem, (cudaStream_t)stream, (const __nv_bfloat16*)dout, p[0], p[1], p[2], w,
                                                                 dp[0], dp[1], dp[2], dw, accumulate, J, HW, C,
                                                                 (long long)lo_delta, mp_deterministic()));
  MP_CHECK_LAUNCH("mp_combiner_bwd");
  return MP_OK;
}

int mp_stem_im2col(const float* x, void* patches, int N, int H, int W, int64_t lo_delta, void* stream) {
  MP_CHECK_DELTA("mp_stem_im2col");
  MP_CHECK_ARG(x && patches && N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0, "mp_stem_im2col: bad arguments");
  const long long blocks = (long long)N * (H / 2) * ((W / 2 + STEM_TW - 1) / STEM_TW);
  MP_CHECK_ARG(blocks <= 0x7fffffffLL, "mp_stem_im2col: too many tiles");
  MP_CUDA(mp_launch(stem_im2col_tiled_kernel<false>, dim3((unsigned)blocks), dim3(192), 0, (cudaStream_t)stream,
      (const void*)x, (__nv_bfloat16*)patches, N, H, W, make_float3(1.f, 1.f, 1.f), make_float3(0.f, 0.f, 0.f),
      (long long)lo_delta));
  MP_CHECK_LAUNCH("mp_stem_im2col");
  return MP_OK;
}

int mp_stem_im2col_u8(const uint8_t* x, void* patches, const float mean[3], const float stddev[3], int N, int H, int W,
                      int64_t lo_delta, void* stream) {
  MP_CHECK_DELTA("mp_stem_im2col_u8");
  MP_CHECK_ARG(x && patches && mean && stddev && N > 0 && H > 0 && W > 0 && H % 2 == 0 && W % 2 == 0,
               "mp_stem_im2col_u8: bad arguments");
  MP_CHECK_ARG(stddev[0] > 0.f && stddev[1] > 0.f && stddev[2] > 0.f, "mp_stem_im2col_u8: stddev must be positive");
  // (p / 255 - mean) / std = p * scale + shift
  const float3 scale = make_float3(1.f / (255.f * stddev[0]), 1.f / (255.f * stddev[1]), 1.f / (255.f * stddev[2]));
  const float3 shift = make_float3(-mean[0] / stddev[0], -mean[1] / stddev[1], -mean[2] / stddev[2]);
  const long long blocks = (long long)N * (H / 2) * ((W / 2 + STEM_TW - 1) / STEM_TW);
  MP_CHECK_ARG(blocks <= 0x7fffffffLL, "mp_stem_im2col_u8: too many tiles");
  MP_CUDA(mp_launch(stem_im2col_tiled_kernel<true>, dim3((unsigned)blocks), dim3(192), 0, (cudaStream_t)stream,
      (const void*)x, (__nv_bfloat16*)patches, N, H, W, scale, shift, (long long)lo_delta));
  MP_CHECK_LAUNCH("mp_stem_im2col_u8");
  return MP_OK;
}

int mp_add_bf16(const void* const in[4], int n, void* out, int64_t count, int64_t lo_delta, void* stream) {
  MP_CHECK_DELTA("mp_add_bf16");
  MP_CHECK_ARG(in && out && n >= 1 && n <= 4 && count > 0 && count % 8 == 0, "mp_add_bf16: bad arguments");
  for (int i = 0; i < n; ++i) MP_CHECK_ARG(in[i] && mp_aligned16(in[i]), "mp_add_bf16: bad input %d", i);
  const long long groups = count / 8;
  MP_CUDA(mp_launch(add_bf16_kernel, dim3((unsigned)((groups + 255) / 256)), dim3(256), 0, (cudaStream_t)stream, 
      (const __nv_bfloat16*)in[0], (const __nv_bfloat16*)(n > 1 ? in[1] : nullptr),
      (const __nv_bfloat16*)(n > 2 ? in[2] : nullptr), (const __nv_bfloat16*)(n > 3 ? in[3] : nullptr), n,
      (__nv_bfloat16*)out, groups, (long long)lo_delta));
  MP_CHECK_LAUNCH("mp_add_bf16");
  return MP_OK;
}

int mp_pack_weights(const float* master, void* packed, const mp_pack_entry* table, int n_entries,
                    int64_t total_work, void* stream) {
  MP_CHECK_ARG(master && packed && table && n_entries > 0 && total_work > 0 && total_work % 8 == 0,
               "mp_pack_weights: bad arguments");
  const long long groups = total_work / 8;
  pack_weights_kernel<<<(unsigned)((groups + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      master, (__nv_bfloat16*)packed, table, n_entries, groups);
  MP_CHECK_LAUNCH("mp_pack_weights");
  return MP_OK;
}

int mp_sgd_step(float* param, const float* grad, float* momentum_buf, int64_t n, float lr, float momentum,
                float dampening, float weight_decay, int nesterov, int first_step, float grad_scale,
                void* stream) {
  return mp_sgd_step_hp(param, grad, momentum_buf, n, lr, momentum, dampening, weight_decay, nesterov, first_step,
                        grad_scale, nullptr, stream);
}

int mp_sgd_step_hp(float* param, const float* grad, float* momentum_buf, int64_t n, float lr, float momentum,
                   float dampening, float weight_decay, int nesterov, int first_step, float grad_scale,
                   const float* hyper, void* stream) {
  MP_CHECK_ARG(param && grad && n > 0 && ((momentum == 0.f && !hyper) || momentum_buf), "mp_sgd_step: bad arguments");
  MP_CHECK_ARG(((reinterpret_cast<uintptr_t>(param) | reinterpret_cast<uintptr_t>(grad) |
                 reinterpret_cast<uintptr_t>(momentum_buf)) & 15) == 0, "mp_sgd_step: buffers must be 16-byte aligned");
  long long blocks = ((n >> 2) + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  if (blocks < 1) blocks = 1;
  sgd_kernel<<<(unsigned)blocks, 256, 0, (cudaStream_t)stream>>>(
      param, grad, momentum_buf, n, lr, momentum, dampening, weight_decay, nesterov, first_step, grad_scale, hyper);
  MP_CHECK_LAUNCH("mp_sgd_step");
  return MP_OK;
}

}  // extern "C"
