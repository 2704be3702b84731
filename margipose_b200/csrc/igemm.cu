// Convolution forward / data-gradient as an implicit GEMM on tcgen05 (sm_100a).
//
//   D[pixel, n] = sum_taps sum_c  A_tap[pixel, c] * W[n, koff_tap + c]
//
// A_tap is never materialised: one TMA box {64 ch, tile_w, 1, rows, 1} of the 5-D activation
// view lands a shifted (and, for stride 2, decimated) pixel x 64-channel operand tile in shared
// memory in the 128B-swizzled K-major layout tcgen05.mma reads; image borders are TMA
// out-of-bounds zero fill.  Taps that differ only by their row shift dh (the three rows of a 3x3
// filter column) form a TAP GROUP: their operand tiles are row-shifted windows of ONE box of
// tile_rows + 2 rows, so the box is fetched once and each tap's MMA descriptor starts
// dh * tile_w * 128 bytes further (a whole number of 1024-byte swizzle atoms) -- the kernel is
// bound by L2 -> shared-memory delivery (profiles/), and this halves the activation traffic.
// W tiles ([n_tile rows] x 64 K) arrive through their own ring.  One elected thread issues
// tcgen05.mma (M=128, N=n_tile, K=16) into a TMEM accumulator; four epilogue warps pull it back
// with tcgen05.ld, add the optional residual, round to bf16, store NHWC and reduce the BatchNorm
// batch statistics of the stored tile.
//
// CTA-pair mode (cta_group::2): two CTAs of a cluster own adjacent 128-pixel tiles and each loads
// only HALF of every weight tile; the leader issues M=256 MMAs that read both halves, so the
// weight traffic per SM halves too.
//
// Replaces the cuDNN calls behind nn.Conv2d / nn.ConvTranspose2d on the reference's hot path
// (/root/reference/src/margipose/models/margipose_model.py:33,67-68,73-74,79-82 and the
// torchvision ResNet stem :130-135), forward and dgrad.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2-5 = epilogue (TMEM lane quarter = warp % 4).  Roofline: tensor pipe; algorithmic
// flops per launch = 2 * pixels * n * K.
#include <string.h>
#include "tc.cuh"
#include "../../include/margipose_b200.h"

namespace {

constexpr int NTHREADS = 192;
constexpr int A_TILE_BYTES = 128 * 128;   // 128 pixels x 64 bf16
constexpr int MAX_STAGES = 8;

// taps (src, c0, dw, p) with row shifts dh0 .. dh0 + n - 1; n > 1 reads the halo box (tile_rows + 2 rows)
struct TapGroup {
  int src, c0, dw, p, dh0, n;
  int koff[3];
};

struct alignas(64) IgemmMaps {   // TMA descriptors per problem: activation sources (plain / halo box) and weights
  CUtensorMap a0[MP_MAX_GROUP], a0h[MP_MAX_GROUP], a1[MP_MAX_GROUP], b[MP_MAX_GROUP];
};

struct IgemmParams {
  TapGroup groups[MP_MAX_TAPS];
  int n_groups, cblocks;
  int tile_w, tile_rows, tiles_w, tiles_h;
  int out_h, out_w;
  int n_tile, tmem_cols;
  int mt;        // 128-pixel accumulators per CTA (1 or 2): row blocks h0 + i * tile_rows share every weight tile
  int a_stages, b_stages, a_slot_bytes, b_slot_bytes;   // b_slot_bytes: this CTA's share of a weight tile
  int a_tx_plain, a_tx_halo, row_bytes;
  int dbg;       // experiment switches (tunable igemm_dbg): 1 = no TMA loads, 2 = no MMA, 4 = no epilogue
  int pair;      // 1: CTA pair (cluster of 2 along the M tiles), cta_group::2 MMA
  long long out_sn, out_sh, out_sw;
  int out_c;
  int stat_replicas;
  long long stat_stride;
  // per problem of a grouped launch (blockIdx.z): same geometry, different tensors
  struct Problem {
    __nv_bfloat16* out;
    const __nv_bfloat16* res;
    float* stat_sum;
    float* stat_sq;
    // BatchNorm finalize by the last CTA (see mp_igemm_args.bn)
    const float* fin_gamma; const float* fin_beta; const float* fin_bias;
    float* fin_rmean; float* fin_rvar; float* fin_smean; float* fin_sinvstd; float* fin_scale; float* fin_shift;
    unsigned* fin_counter;
  } q[MP_MAX_GROUP];
  int fin_total, fin_C, fin_Cp;
  long long fin_count;
  float fin_momentum, fin_eps;
};

template <bool PAIR>
__global__ void __launch_bounds__(NTHREADS)
igemm_kernel(const __grid_constant__ IgemmMaps TM, const __grid_constant__ IgemmParams P) {
  const CUtensorMap& tmA0 = TM.a0[blockIdx.z];
  const CUtensorMap& tmA0h = TM.a0h[blockIdx.z];
  const CUtensorMap& tmA1 = TM.a1[blockIdx.z];
  const CUtensorMap& tmB = TM.b[blockIdx.z];
  const IgemmParams::Problem& Q = P.q[blockIdx.z];
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)P.a_stages * P.a_slot_bytes;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(sB + (size_t)P.b_stages * P.b_slot_bytes);
  uint64_t* emptyA = fullA + MAX_STAGES;
  uint64_t* fullB = emptyA + MAX_STAGES;
  uint64_t* emptyB = fullB + MAX_STAGES;
  uint64_t* tmem_full = emptyB + MAX_STAGES;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* s_stat = reinterpret_cast<float*>(tmem_holder + 4);   // [2][256] per-CTA channel sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (Q.stat_sum)
    for (int i = threadIdx.x; i < 512; i += NTHREADS) s_stat[i] = 0.f;
  int t = blockIdx.x;
  const int tw = t % P.tiles_w; t /= P.tiles_w;
  const int th = t % P.tiles_h;
  const int img = t / P.tiles_h;
  const int w0 = tw * P.tile_w, h0 = th * P.tile_rows * P.mt;
  const int n0 = blockIdx.y * P.n_tile;
  const uint32_t crank = PAIR ? tc::cluster_ctarank() : 0;

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      tc::mbar_init(&fullA[s], 1);
      tc::mbar_init(&emptyA[s], 1);
      tc::mbar_init(&fullB[s], 1);
      tc::mbar_init(&emptyB[s], 1);
    }
    tc::mbar_init(tmem_full, 1);
    tc::mbar_fence_init();
    tc::prefetch_tmap(&tmA0);
    tc::prefetch_tmap(&tmA0h);
    tc::prefetch_tmap(&tmA1);
    tc::prefetch_tmap(&tmB);
  }
  if (warp == 1) {
    if (PAIR) tc::tmem_alloc_pair(tmem_holder, P.tmem_cols);
    else tc::tmem_alloc(tmem_holder, P.tmem_cols);
  }
  tc::tc_fence_before();
  __syncthreads();
  if (PAIR) tc::cluster_sync_all();   // the peer's barriers are initialised before anyone signals them
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {   // ---------------------------------------------------------- TMA producer
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      const int b_row0 = n0 + (int)crank * (P.b_slot_bytes >> 7);   // pair: my half of the weight-tile rows
      for (int g = 0; g < P.n_groups; ++g) {
        const TapGroup& G = P.groups[g];
        const bool halo = G.n > 1;
        const CUtensorMap* tmA = halo ? &tmA0h : (G.src ? &tmA1 : &tmA0);
        for (int cb = 0; cb < P.cblocks; ++cb) {
          tc::mbar_wait(&emptyA[sa], pha ^ 1);
          uint8_t* dstA = sA + (size_t)sa * P.a_slot_bytes;
          if (P.dbg & 1) {
            if (crank == 0) tc::mbar_arrive(&fullA[sa]);
          } else if (PAIR) {   // both CTAs' bytes complete on the leader's barrier
            if (crank == 0) tc::mbar_arrive_expect_tx(&fullA[sa], 2u * (uint32_t)(halo ? P.a_tx_halo : P.a_tx_plain));
            tc::tma_load_5d_pair(tmA, &fullA[sa], dstA, G.c0 + cb * 64, w0 + G.dw, G.p, h0 + G.dh0, img);
          } else {
            tc::mbar_arrive_expect_tx(&fullA[sa], (uint32_t)(halo ? P.a_tx_halo : P.a_tx_plain));
            tc::tma_load_5d(tmA, &fullA[sa], dstA, G.c0 + cb * 64, w0 + G.dw, G.p, h0 + G.dh0, img);
          }
          if (++sa == P.a_stages) { sa = 0; pha ^= 1; }
          for (int i = 0; i < G.n; ++i) {
            tc::mbar_wait(&emptyB[sb], phb ^ 1);
            uint8_t* dstB = sB + (size_t)sb * P.b_slot_bytes;
            if (P.dbg & 1) {
              if (crank == 0) tc::mbar_arrive(&fullB[sb]);
            } else if (PAIR) {
              if (crank == 0) tc::mbar_arrive_expect_tx(&fullB[sb], 2u * (uint32_t)P.b_slot_bytes);
              tc::tma_load_2d_pair(&tmB, &fullB[sb], dstB, G.koff[i] + cb * 64, b_row0);
            } else {
              tc::mbar_arrive_expect_tx(&fullB[sb], (uint32_t)P.b_slot_bytes);
              tc::tma_load_2d(&tmB, &fullB[sb], dstB, G.koff[i] + cb * 64, b_row0);
            }
            if (++sb == P.b_stages) { sb = 0; phb ^= 1; }
          }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0 && crank == 0) {   // -------------------------------- MMA issuer (pair: the leader only)
      const uint32_t idesc = tc::idesc_bf16(PAIR ? 256 : 128, P.n_tile, false, false);
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      bool first = true;
      for (int g = 0; g < P.n_groups; ++g) {
        const int n = P.groups[g].n;
        for (int cb = 0; cb < P.cblocks; ++cb) {
          tc::mbar_wait(&fullA[sa], pha);
          const uint32_t a0 = tc::smem_u32(sA + (size_t)sa * P.a_slot_bytes);
          for (int i = 0; i < n; ++i) {
            tc::mbar_wait(&fullB[sb], phb);
            tc::tc_fence_after();
            const uint64_t bd = tc::desc_kmajor_sw128(tc::smem_u32(sB + (size_t)sb * P.b_slot_bytes));
            for (int m = 0; m < P.mt; ++m) {
              // row-shifted window of the (halo) box: tap i of the group, row block m of the CTA
              const uint64_t ad = tc::desc_kmajor_sw128(a0 + (uint32_t)((i + m * P.tile_rows) * P.row_bytes));
              const uint32_t d = tmem + (uint32_t)(m * P.n_tile);
#pragma unroll
              for (int k = 0; k < 4; ++k) {   // 4 x (K = 16 bf16 = 32 bytes inside the 128B swizzle atom: +2 in the
                                              // descriptor's 16-byte address field)
                if (P.dbg & 2) continue;
                if (PAIR) tc::mma_bf16_pair(d, ad + 2 * k, bd + 2 * k, idesc, !(first && k == 0));
                else tc::mma_bf16(d, ad + 2 * k, bd + 2 * k, idesc, !(first && k == 0));
              }
            }
            first = false;
            // frees the weight slot (in both CTAs of a pair) once these MMAs have read it
            if (PAIR) tc::mma_commit_pair(&emptyB[sb]); else tc::mma_commit(&emptyB[sb]);
            if (++sb == P.b_stages) { sb = 0; phb ^= 1; }
          }
          if (PAIR) tc::mma_commit_pair(&emptyA[sa]); else tc::mma_commit(&emptyA[sa]);
          if (++sa == P.a_stages) { sa = 0; pha ^= 1; }
        }
      }
      if (PAIR) tc::mma_commit_pair(tmem_full); else tc::mma_commit(tmem_full);
    }
  } else {   // ------------------------------------------------------------------------ epilogue
    tc::mbar_wait(tmem_full, 0);
    tc::tc_fence_after();
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int r = m / P.tile_w, wq = m - r * P.tile_w;
    const int nchunks = P.n_tile / 32;
    for (int mt = 0; mt < ((P.dbg & 4) ? 0 : P.mt); ++mt)
    for (int c = 0; c < nchunks; ++c) {
      const int h = h0 + mt * P.tile_rows + r, w = w0 + wq;
      const bool valid = (r < P.tile_rows) && (h < P.out_h) && (w < P.out_w);
      const long long pix = (long long)img * P.out_sn + (long long)h * P.out_sh + (long long)w * P.out_sw;
      const int ch = n0 + c * 32;
      if (ch >= P.out_c) break;   // warp-uniform
      float v[32];
      tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * P.n_tile + c * 32), v);
      uint32_t packed[16];
      if (valid) {
        if (Q.res) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (ch + i * 8 < P.out_c) {
              const uint4 rv = __ldg(reinterpret_cast<const uint4*>(Q.res + pix + ch + i * 8));
              const uint32_t rr[4] = {rv.x, rv.y, rv.z, rv.w};
#pragma unroll
              for (int j = 0; j < 4; ++j) {
                const float2 f = unpack_bf16x2(rr[j]);
                v[i * 8 + j * 2] += f.x;
                v[i * 8 + j * 2 + 1] += f.y;
              }
            }
          }
        }
#pragma unroll
        for (int j = 0; j < 16; ++j) packed[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          if (ch + i * 8 < P.out_c)
            *reinterpret_cast<uint4*>(Q.out + pix + ch + i * 8) =
                make_uint4(packed[i * 4], packed[i * 4 + 1], packed[i * 4 + 2], packed[i * 4 + 3]);
        }
      }
      if (Q.stat_sum) {   // statistics of the values as stored (bf16-rounded)
        float s1[32], s2[32];
#pragma unroll
        for (int j = 0; j < 16; ++j) {
          float2 f = make_float2(0.f, 0.f);
          if (valid) f = unpack_bf16x2(packed[j]);
          s1[2 * j] = f.x; s1[2 * j + 1] = f.y;
          s2[2 * j] = f.x * f.x; s2[2 * j + 1] = f.y * f.y;
        }
        const float a1 = tc::warp_transpose_sum(s1);
        const float a2 = tc::warp_transpose_sum(s2);
        atomicAdd(s_stat + c * 32 + lane, a1);          // 4 epilogue warps meet in shared memory ...
        atomicAdd(s_stat + 256 + c * 32 + lane, a2);
      }
    }
    if (Q.stat_sum) {   // ... and the CTA issues ONE global atomic per channel and statistic
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const long long rep = (long long)(blockIdx.x % P.stat_replicas) * P.stat_stride;
      for (int i = threadIdx.x - 64; i < P.n_tile; i += 128) {
        if (n0 + i < P.out_c) {
          atomicAdd(Q.stat_sum + rep + n0 + i, s_stat[i]);
          atomicAdd(Q.stat_sq + rep + n0 + i, s_stat[256 + i]);
        }
      }
      if (Q.fin_counter) {   // last CTA to arrive turns the sums into the BatchNorm coefficients
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        unsigned* s_ticket = reinterpret_cast<unsigned*>(s_stat + 512);
        if (threadIdx.x == 64) *s_ticket = atomicAdd(Q.fin_counter, 1u);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (*s_ticket == (unsigned)(P.fin_total - 1)) {
          __threadfence();
          const float inv_m = 1.0f / (float)P.fin_count;
          for (int c = threadIdx.x - 64; c < P.fin_Cp; c += 128) {
            float scale = 0.f, shift = 0.f;
            if (c < P.fin_C) {
              const float mean = __ldcg(Q.stat_sum + c) * inv_m;
              const float var = fmaxf(__ldcg(Q.stat_sq + c) * inv_m - mean * mean, 0.f);
              const float invstd = rsqrtf(var + P.fin_eps);
              scale = Q.fin_gamma[c] * invstd;
              shift = fmaf(-mean, scale, Q.fin_beta[c]);
              Q.fin_smean[c] = mean;
              Q.fin_sinvstd[c] = invstd;
              if (Q.fin_rmean) {
                const float unbiased = P.fin_count > 1 ? var * ((float)P.fin_count / (float)(P.fin_count - 1)) : var;
                const float bias = Q.fin_bias ? Q.fin_bias[c] : 0.f;
                Q.fin_rmean[c] = (1.f - P.fin_momentum) * Q.fin_rmean[c] + P.fin_momentum * (mean + bias);
                Q.fin_rvar[c] = (1.f - P.fin_momentum) * Q.fin_rvar[c] + P.fin_momentum * unbiased;
              }
            }
            Q.fin_scale[c] = scale;
            Q.fin_shift[c] = shift;
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (PAIR) {
    tc::cluster_sync_all();   // nobody leaves while the pair's MMAs may still read its shared memory / TMEM
    if (warp == 1) tc::tmem_dealloc_pair(tmem, P.tmem_cols);
  } else if (warp == 1) {
    tc::tmem_dealloc(tmem, P.tmem_cols);
  }
}

long long g_igemm_smem = 115712;   // 113 KB: two CTAs share an SM (one's epilogue overlaps the other's main loop)
long long g_igemm_smem2 = 200 * 1024;   // budget of CTAs with two accumulators (one per SM, deep rings)
long long g_igemm_mt = 1;          // row blocks (accumulators) per CTA when the launch keeps >= g_igemm_mt_ctas CTAs
long long g_igemm_mt_ctas = 100;
long long g_igemm_halo = 1;        // group taps that differ only in their row shift (one halo box per group)
long long g_igemm_dbg = 0;
long long g_igemm_pair = 0;        // CTA pairs (cta_group::2) when the M tiles pair up
long long g_igemm_split_n = 100;   // split N in two when the launch would have fewer CTAs than this
bool g_attr_set = false;

int view_to_tmap(CUtensorMap* tm, const mp_view5& v, const uint32_t box[5], const char* what) {
  uint64_t dims[5], strides[5];
  for (int i = 0; i < 5; ++i) {
    MP_CHECK_ARG(v.dim[i] > 0, "%s: view dim %d is %lld", what, i, (long long)v.dim[i]);
    dims[i] = (uint64_t)v.dim[i];
    strides[i] = (uint64_t)v.stride[i] * 2;
    MP_CHECK_ARG(i == 0 || strides[i] % 16 == 0, "%s: view stride %d not 16-byte aligned", what, i);
  }
  MP_CHECK_ARG(v.stride[0] == 1, "%s: innermost view stride must be 1", what);
  MP_CHECK_ARG(mp_aligned16(v.ptr), "%s: view pointer not 16-byte aligned", what);
  return tc::encode_tmap(tm, v.ptr, 5, dims, strides, box);
}

// Partition the taps into groups of up to three that share (src 0, c0, dw, p) and have consecutive
// row shifts; everything else becomes a group of one.  Returns the number of groups.
int group_taps(const mp_igemm_args* a, bool allow_halo, TapGroup* out) {
  bool used[MP_MAX_TAPS] = {};
  int n_groups = 0;
  for (int i = 0; i < a->n_taps; ++i) {
    if (used[i]) continue;
    const mp_tap& t = a->taps[i];
    used[i] = true;
    int dh[3] = {t.dh, 0, 0}, koff[3] = {t.koff, 0, 0}, n = 1;
    int lo = t.dh, hi = t.dh;
    if (allow_halo && t.src == 0) {
      bool grew = true;
      while (grew && n < 3) {
        grew = false;
        for (int j = 0; j < a->n_taps && n < 3; ++j) {
          const mp_tap& u = a->taps[j];
          if (used[j] || u.src != 0 || u.c0 != t.c0 || u.dw != t.dw || u.p != t.p) continue;
          if (u.dh != hi + 1 && u.dh != lo - 1) continue;
          if (u.dh > hi) hi = u.dh; else lo = u.dh;
          dh[n] = u.dh; koff[n] = u.koff; ++n;
          used[j] = true;
          grew = true;
        }
      }
    }
    TapGroup G;
    G.src = t.src; G.c0 = t.c0; G.dw = t.dw; G.p = t.p; G.dh0 = lo; G.n = n;
    G.koff[0] = G.koff[1] = G.koff[2] = 0;
    for (int k = 0; k < n; ++k) G.koff[dh[k] - lo] = koff[k];
    out[n_groups++] = G;
  }
  return n_groups;
}

}  // namespace

int mp_pick_tile(int out_h, int out_w, int max_pix, int* tile_w, int* tile_rows) {
  int tw = out_w;
  if (tw > max_pix) {
    tw = max_pix;
    for (int d = max_pix; d >= max_pix / 2; --d)
      if (out_w % d == 0) { tw = d; break; }
  }
  int tr = max_pix / tw;
  if (tr > out_h) tr = out_h;
  if (tr < 1) tr = 1;
  *tile_w = tw;
  *tile_rows = tr;
  return 0;
}

void mp_set_igemm_smem(long long v) { g_igemm_smem = v; }
void mp_set_igemm_halo(long long v) { g_igemm_halo = v; }
void mp_set_igemm_pair(long long v) { g_igemm_pair = v; }
void mp_set_igemm_dbg(long long v) { g_igemm_dbg = v; }
void mp_set_igemm_split_n(long long v) { g_igemm_split_n = v; }
void mp_set_igemm_mt(int which, long long v) {
  if (which == 0) g_igemm_mt = v;
  else if (which == 1) g_igemm_mt_ctas = v;
  else g_igemm_smem2 = v;
}

static void igemm_grid(const mp_igemm_args* a, int* tile_w, int* tile_rows, int* tiles_w, int* tiles_h, int* n_tile,
                       int* mt) {
  mp_pick_tile(a->out_h, a->out_w, 128, tile_w, tile_rows);
  *tiles_w = (a->out_w + *tile_w - 1) / *tile_w;
  *n_tile = a->w_rows <= 256 ? (int)a->w_rows : 256;
  // Two full 128-pixel row blocks per CTA (two TMEM accumulators sharing every weight tile) when that
  // still leaves enough CTAs: halves the weight traffic per FLOP and amortises the MMA issuer's
  // barrier handshakes over twice the math.
  *mt = 1;
  if (g_igemm_mt >= 2 && *tile_w * *tile_rows == 128 && a->out_h >= 2 * *tile_rows) {
    const long long th2 = (a->out_h + 2 * *tile_rows - 1) / (2 * *tile_rows);
    if ((long long)a->n_img * th2 * *tiles_w * (a->w_rows / *n_tile) >= g_igemm_mt_ctas) *mt = 2;
  }
  *tiles_h = (a->out_h + *tile_rows * *mt - 1) / (*tile_rows * *mt);
  // Small pixel grids (e.g. 16x16 maps at batch 32 = 64 M-tiles) leave most of the 148 SMs idle:
  // give each M-tile two CTAs with half the output channels each (the A tile is then fetched twice,
  // from L2, which is cheaper than idle tensor cores).
  const long long ctas = (long long)a->n_img * *tiles_h * *tiles_w * (a->w_rows / *n_tile);
  if (ctas < g_igemm_split_n && *n_tile >= 128 && (*n_tile / 2) % 32 == 0) *n_tile /= 2;
}

extern "C" int mp_conv_igemm_ctas(const mp_igemm_args* a) {
  if (!a || a->out_h <= 0 || a->out_w <= 0 || a->w_rows <= 0) return 0;
  int tw, tr, tsw, tsh, nt, mt;
  igemm_grid(a, &tw, &tr, &tsw, &tsh, &nt, &mt);
  return a->n_img * tsh * tsw * (int)(a->w_rows / nt);
}

static int check_one(const mp_igemm_args* a, bool* use_src1) {
  MP_CHECK_ARG(a->n_taps >= 1 && a->n_taps <= MP_MAX_TAPS, "mp_conv_igemm: n_taps %d out of range", a->n_taps);
  MP_CHECK_ARG(a->cblocks >= 1, "mp_conv_igemm: cblocks must be >= 1");
  MP_CHECK_ARG(a->wmat && a->out && a->src[0].ptr, "mp_conv_igemm: null tensor");
  MP_CHECK_ARG(a->w_rows >= 32 && a->w_rows % 32 == 0, "mp_conv_igemm: w_rows %lld must be a multiple of 32",
               (long long)a->w_rows);
  MP_CHECK_ARG(a->w_k % 64 == 0, "mp_conv_igemm: w_k %lld must be a multiple of 64", (long long)a->w_k);
  MP_CHECK_ARG(a->n_img > 0 && a->out_h > 0 && a->out_w > 0, "mp_conv_igemm: empty output grid");
  MP_CHECK_ARG(a->out_c > 0 && a->out_c % 8 == 0 && a->out_c <= a->w_rows,
               "mp_conv_igemm: out_c %d must be a multiple of 8 and <= w_rows", a->out_c);
  MP_CHECK_ARG(mp_aligned16(a->out) && (!a->res || mp_aligned16(a->res)) && a->out_sn % 8 == 0 &&
                   a->out_sh % 8 == 0 && a->out_sw % 8 == 0,
               "mp_conv_igemm: output addressing must be 16-byte aligned");
  MP_CHECK_ARG((a->stat_sum == nullptr) == (a->stat_sq == nullptr), "mp_conv_igemm: stat_sum/stat_sq go together");
  *use_src1 = false;
  for (int i = 0; i < a->n_taps; ++i) {
    MP_CHECK_ARG(a->taps[i].src == 0 || a->taps[i].src == 1, "mp_conv_igemm: tap %d: bad src", i);
    MP_CHECK_ARG(a->taps[i].koff >= 0 && a->taps[i].koff + a->cblocks * 64 <= a->w_k,
                 "mp_conv_igemm: tap %d: K range outside the weight matrix", i);
    if (a->taps[i].src == 1) *use_src1 = true;
  }
  MP_CHECK_ARG(!*use_src1 || a->src[1].ptr, "mp_conv_igemm: tap refers to a missing second source");
  MP_CHECK_ARG(mp_aligned16(a->wmat), "mp_conv_igemm: wmat not 16-byte aligned");
  return MP_OK;
}

static bool same_view(const mp_view5& x, const mp_view5& y) {
  for (int i = 0; i < 5; ++i)
    if (x.dim[i] != y.dim[i] || x.stride[i] != y.stride[i]) return false;
  return true;
}

// Problems of a grouped launch must differ in their tensors only.
static bool same_geometry(const mp_igemm_args* x, const mp_igemm_args* y) {
  if (x->n_taps != y->n_taps || x->cblocks != y->cblocks || x->w_rows != y->w_rows || x->w_k != y->w_k ||
      x->n_img != y->n_img || x->out_h != y->out_h || x->out_w != y->out_w || x->out_sn != y->out_sn ||
      x->out_sh != y->out_sh || x->out_sw != y->out_sw || x->out_c != y->out_c ||
      (x->res == nullptr) != (y->res == nullptr) || (x->stat_sum == nullptr) != (y->stat_sum == nullptr) ||
      (x->bn == nullptr) != (y->bn == nullptr) || x->stat_replicas != y->stat_replicas ||
      x->stat_stride != y->stat_stride || !same_view(x->src[0], y->src[0]) ||
      (x->src[1].ptr == nullptr) != (y->src[1].ptr == nullptr) || (x->src[1].ptr && !same_view(x->src[1], y->src[1])))
    return false;
  if (x->bn && (x->bn_total_ctas != y->bn_total_ctas || x->bn_channels != y->bn_channels || x->bn_count != y->bn_count ||
                x->bn_momentum != y->bn_momentum || x->bn_eps != y->bn_eps))
    return false;
  return memcmp(x->taps, y->taps, sizeof(mp_tap) * x->n_taps) == 0;
}

extern "C" int mp_conv_igemm(const mp_igemm_args* a, void* stream) { return mp_conv_igemm_grouped(a, 1, stream); }

extern "C" int mp_conv_igemm_grouped(const mp_igemm_args* args, int n_problems, void* stream) {
  MP_CHECK_ARG(args, "mp_conv_igemm: null args");
  MP_CHECK_ARG(n_problems >= 1 && n_problems <= MP_MAX_GROUP, "mp_conv_igemm_grouped: %d problems (1..%d)", n_problems,
               MP_MAX_GROUP);
  const mp_igemm_args* a = &args[0];
  bool use_src1 = false;
  for (int i = 0; i < n_problems; ++i) {
    bool u;
    int rc = check_one(&args[i], &u);
    if (rc != MP_OK) return rc;
    if (i == 0) use_src1 = u;
    MP_CHECK_ARG(i == 0 || same_geometry(a, &args[i]), "mp_conv_igemm_grouped: problem %d differs in geometry", i);
  }

  IgemmParams P;
  igemm_grid(a, &P.tile_w, &P.tile_rows, &P.tiles_w, &P.tiles_h, &P.n_tile, &P.mt);
  P.cblocks = a->cblocks;
  P.out_h = a->out_h;
  P.out_w = a->out_w;
  MP_CHECK_ARG(a->w_rows % P.n_tile == 0, "mp_conv_igemm: w_rows %lld not tileable", (long long)a->w_rows);

  // tap groups: row-shifted taps share one halo box when a pixel row is a whole number of swizzle atoms
  P.row_bytes = P.tile_w * 128;
  const bool allow_halo = g_igemm_halo != 0 && P.tile_w % 8 == 0;
  P.n_groups = group_taps(a, allow_halo, P.groups);
  bool any_halo = false;
  for (int g = 0; g < P.n_groups; ++g) any_halo |= P.groups[g].n > 1;
  P.a_tx_plain = P.tile_w * P.tile_rows * P.mt * 128;
  P.a_tx_halo = P.tile_w * (P.tile_rows * P.mt + 2) * 128;
  P.a_slot_bytes = (P.mt - 1) * P.tile_rows * P.row_bytes + A_TILE_BYTES + (any_halo ? 2 * P.row_bytes : 0);

  // CTA pairs: adjacent M tiles, each CTA holds half of the weight-tile rows
  const long long m_tiles = (long long)a->n_img * P.tiles_h * P.tiles_w;
  P.pair = (g_igemm_pair != 0 && m_tiles % 2 == 0 && P.n_tile % 16 == 0 && (P.n_tile / 2) % 8 == 0) ? 1 : 0;
  P.b_slot_bytes = (P.n_tile / (P.pair ? 2 : 1)) * 128;
  P.dbg = (int)g_igemm_dbg;

  // shared-memory rings: with halo groups each activation box feeds up to three weight tiles
  const int overhead = 1024 + 512 + 2048 + 64;   // alignment slack + barriers + per-CTA channel sums + ticket
  const long long budget = (P.mt > 1 ? g_igemm_smem2 : g_igemm_smem) - overhead;
  const int a_loads = P.n_groups * a->cblocks, b_loads = a->n_taps * a->cblocks;
  int a_stages, b_stages;
  if (any_halo) {
    a_stages = 2;
    b_stages = (int)((budget - (long long)a_stages * P.a_slot_bytes) / P.b_slot_bytes);
    if (b_stages > 6 && budget - 3LL * P.a_slot_bytes - 6LL * P.b_slot_bytes >= 0) a_stages = 3;
    b_stages = (int)((budget - (long long)a_stages * P.a_slot_bytes) / P.b_slot_bytes);
  } else {
    a_stages = b_stages = (int)(budget / (P.a_slot_bytes + P.b_slot_bytes));
  }
  if (a_stages < 2) a_stages = 2;
  if (b_stages < 2) b_stages = 2;
  if (a_stages > MAX_STAGES) a_stages = MAX_STAGES;
  if (b_stages > MAX_STAGES) b_stages = MAX_STAGES;
  if (a_stages > a_loads) a_stages = a_loads < 2 ? 2 : a_loads;
  if (b_stages > b_loads) b_stages = b_loads < 2 ? 2 : b_loads;
  P.a_stages = a_stages;
  P.b_stages = b_stages;
  {
    const int cols = P.n_tile * P.mt;
    P.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
  }
  P.out_sn = a->out_sn; P.out_sh = a->out_sh; P.out_sw = a->out_sw;
  P.out_c = a->out_c;
  P.stat_replicas = a->stat_replicas > 1 ? a->stat_replicas : 1;
  P.stat_stride = a->stat_stride;
  P.fin_total = 0; P.fin_C = 0; P.fin_Cp = 0; P.fin_count = 0; P.fin_momentum = 0.f; P.fin_eps = 0.f;
  if (a->bn) {
    MP_CHECK_ARG(a->bn_count > 0 && a->bn_channels > 0 && a->bn_channels <= a->out_c,
                 "mp_conv_igemm: incomplete BatchNorm finalize arguments");
    P.fin_C = a->bn_channels; P.fin_Cp = a->out_c;
    P.fin_count = a->bn_count; P.fin_momentum = a->bn_momentum; P.fin_eps = a->bn_eps;
    P.fin_total = a->bn_total_ctas > 0 ? a->bn_total_ctas
                                       : a->n_img * P.tiles_h * P.tiles_w * (int)(a->w_rows / P.n_tile);
  }

  IgemmMaps TM;
  const uint32_t boxA[5] = {64, (uint32_t)P.tile_w, 1, (uint32_t)(P.tile_rows * P.mt), 1};
  const uint32_t boxH[5] = {64, (uint32_t)P.tile_w, 1, (uint32_t)(P.tile_rows * P.mt + 2), 1};
  for (int i = 0; i < MP_MAX_GROUP; ++i) {
    const mp_igemm_args* x = &args[i < n_problems ? i : 0];
    IgemmParams::Problem& Q = P.q[i];
    Q.out = reinterpret_cast<__nv_bfloat16*>(x->out);
    Q.res = reinterpret_cast<const __nv_bfloat16*>(x->res);
    Q.stat_sum = x->stat_sum;
    Q.stat_sq = x->stat_sq;
    Q.fin_gamma = Q.fin_beta = Q.fin_bias = nullptr;
    Q.fin_rmean = Q.fin_rvar = Q.fin_smean = Q.fin_sinvstd = Q.fin_scale = Q.fin_shift = nullptr;
    Q.fin_counter = nullptr;
    if (x->bn) {
      const mp_bn_branch* b = x->bn;
      MP_CHECK_ARG(x->stat_sum && P.stat_replicas == 1, "mp_conv_igemm: BatchNorm finalize needs un-replicated statistics");
      MP_CHECK_ARG(x->bn_counter && b->gamma && b->beta && b->save_mean && b->save_invstd && b->scale && b->shift,
                   "mp_conv_igemm: incomplete BatchNorm finalize arguments");
      Q.fin_gamma = b->gamma; Q.fin_beta = b->beta; Q.fin_bias = b->conv_bias;
      Q.fin_rmean = b->running_mean; Q.fin_rvar = b->running_var;
      Q.fin_smean = b->save_mean; Q.fin_sinvstd = b->save_invstd; Q.fin_scale = b->scale; Q.fin_shift = b->shift;
      Q.fin_counter = x->bn_counter;
    }
    int rc = view_to_tmap(&TM.a0[i], x->src[0], boxA, "mp_conv_igemm src[0]");
    if (rc != MP_OK) return rc;
    if (any_halo) {
      rc = view_to_tmap(&TM.a0h[i], x->src[0], boxH, "mp_conv_igemm src[0] (halo box)");
      if (rc != MP_OK) return rc;
    } else {
      TM.a0h[i] = TM.a0[i];
    }
    if (use_src1) {
      rc = view_to_tmap(&TM.a1[i], x->src[1], boxA, "mp_conv_igemm src[1]");
      if (rc != MP_OK) return rc;
    } else {
      TM.a1[i] = TM.a0[i];
    }
    const uint64_t dims[2] = {(uint64_t)x->w_k, (uint64_t)x->w_rows};
    const uint64_t strides[2] = {2, (uint64_t)x->w_k * 2};
    const uint32_t box[2] = {64, (uint32_t)(P.b_slot_bytes / 128)};
    rc = tc::encode_tmap(&TM.b[i], x->wmat, 2, dims, strides, box);
    if (rc != MP_OK) return rc;
  }

  const size_t smem = (size_t)a_stages * P.a_slot_bytes + (size_t)b_stages * P.b_slot_bytes + overhead;
  MP_CHECK_ARG(smem <= 227 * 1024, "mp_conv_igemm: %zu bytes of shared memory needed", smem);
  if (!g_attr_set) {
    MP_CUDA(cudaFuncSetAttribute(igemm_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    MP_CUDA(cudaFuncSetAttribute(igemm_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    g_attr_set = true;
  }
  dim3 grid((unsigned)m_tiles, (unsigned)(a->w_rows / P.n_tile), (unsigned)n_problems);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = P.pair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  if (P.pair) MP_CUDA(cudaLaunchKernelEx(&cfg, igemm_kernel<true>, TM, P));
  else MP_CUDA(cudaLaunchKernelEx(&cfg, igemm_kernel<false>, TM, P));
  MP_CHECK_LAUNCH("mp_conv_igemm");
  return MP_OK;
}
