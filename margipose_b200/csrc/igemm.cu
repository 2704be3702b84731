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
// The kernel is PERSISTENT: ~one CTA per SM walks a contiguous range of output tiles; the TMA ring
// keeps running across tile boundaries and the TMEM accumulator is double-buffered, so the epilogue
// of tile i (tcgen05.ld, residual, bf16 stores, BatchNorm sums) overlaps the MMAs of tile i + 1 and
// the per-CTA BatchNorm sums reach global memory once per CTA instead of once per tile.  A grouped
// launch (grid z = problem) runs the same layer of the three HeatmapColumns of a stage at once.
//
// CTA-pair mode (cta_group::2): two CTAs of a cluster own adjacent 128-pixel tiles and each loads
// only HALF of every weight tile; the leader issues M=256 MMAs that read both halves, so the
// weight traffic per SM halves too.
//
// Replaces the cuDNN calls behind nn.Conv2d / nn.ConvTranspose2d on the reference's hot path
// (/root/reference/src/margipose/models/margipose_model.py:33,67-68,73-74,79-82 and the
// torchvision ResNet stem :130-135), forward and dgrad.
//
// Warp roles (320 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2-9 = epilogue (TMEM lane quarter = warp % 4; the two warps of a quarter split the 32-column chunks).  Roofline: tensor pipe; algorithmic
// flops per launch = 2 * pixels * n * K.
#include <string.h>
#include "tc.cuh"
#include "../../include/margipose_b200.h"

namespace {

constexpr int NTHREADS = 320;     // warp 0 producer, warp 1 MMA issuer, warps 2-9 epilogue
constexpr int EPI_THREADS = 256;
constexpr int A_TILE_BYTES = 128 * 128;   // 128 pixels x 64 bf16
constexpr int MAX_STAGES = 8;
constexpr int MAX_B = 40;        // weight slots (resident mode: one per tap and 64-channel block)

// taps (src, c0, dw, p) with row shifts dh0 .. dh0 + n - 1; n > 1 reads the halo box (tile_rows + 2 rows)
struct TapGroup {
  int src, c0, dw, p, dh0, n;
  int koff[3];
};

struct alignas(64) IgemmMaps {   // TMA descriptors per problem: activation sources (plain / halo box) and weights
  CUtensorMap a0[MP_MAX_GROUP], a0h[MP_MAX_GROUP], a1[MP_MAX_GROUP], a1h[MP_MAX_GROUP], b[MP_MAX_GROUP];
};

struct IgemmParams {
  TapGroup groups[MP_MAX_TAPS];
  int n_groups, cblocks;
  int tile_w, tile_rows, tiles_w, tiles_h;   // tiles_h counts CTA row blocks of mt * tile_rows rows
  int out_h, out_w;
  int n_tile, m_tiles, units;   // per problem: m_tiles pixel tiles x (w_rows / n_tile); units = tiles (pair: tile pairs)
  int tmem_cols, acc_cols;      // acc_cols = mt * n_tile columns per accumulator buffer (two buffers)
  int mt;        // 128-pixel accumulators per tile (1 or 2): row blocks h0 + i * tile_rows share every weight tile
  int a_stages, b_stages, a_slot_bytes, b_slot_bytes;   // b_slot_bytes: this CTA's share of a weight tile
  int a_tx_plain, a_tx_halo, row_bytes;
  int resident;  // 1: the CTA's whole weight operand stays in shared memory (one slot per tap and channel block,
                 // loaded with the first tile); later tiles stream activations only
  int prefetch_b;   // 1: request the first weight tiles before griddepcontrol.wait (tunable igemm_prefetch_b)
  int dbg;       // experiment switches (tunable igemm_dbg): 1 = no TMA loads, 2 = no MMA, 4 = no epilogue
  unsigned long long* trace;   // tunable igemm_trace: device buffer of 64 timestamps written by CTA (0,0,0), or NULL
  long long out_sn, out_sh, out_sw;
  int out_c;
  int stat_replicas;
  long long stat_stride;
  // per problem of a grouped launch (blockIdx.z): same geometry, different tensors
  struct Problem {
    __nv_bfloat16* out;
    const __nv_bfloat16* res;
    const __nv_bfloat16* acc_in;   // split mode: partial sum of the earlier launches (added before affine / ReLU)
    float* stat_sum;
    float* stat_sq;
    // BatchNorm finalize by the last CTA (see mp_igemm_args.bn)
    const float* fin_gamma; const float* fin_beta; const float* fin_bias;
    float* fin_rmean; float* fin_rvar; float* fin_smean; float* fin_sinvstd; float* fin_scale; float* fin_shift;
    unsigned* fin_counter;
    // per-channel affine on the accumulator (eval-mode BatchNorm folded into the conv), see mp_igemm_args.ep_scale
    const float* ep_scale; const float* ep_shift;
  } q[MP_MAX_GROUP];
  int ep_relu;   // 0 = none, 1 = ReLU before the residual add, 2 = ReLU after it
  long long lo_delta;   // split (bf16x3) mode: out / res / acc_in are value pairs hi + lo, lo at +lo_delta elements
  int fin_total, fin_C, fin_Cp;
  long long fin_count;
  float fin_momentum, fin_eps;
};

__device__ __forceinline__ void trace_at(const IgemmParams& P, int slot) {
  if (P.trace && blockIdx.x == 0 && blockIdx.z == 0 && slot < 64) {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    P.trace[slot] = t;
  }
}

struct TileCoord {
  int n0, img, h0, w0, n_idx;
};
__device__ __forceinline__ TileCoord decode_tile(const IgemmParams& P, int t) {
  TileCoord c;
  c.n_idx = t / P.m_tiles;
  int m = t - c.n_idx * P.m_tiles;
  const int tw = m % P.tiles_w; m /= P.tiles_w;
  const int th = m % P.tiles_h;
  c.img = m / P.tiles_h;
  c.w0 = tw * P.tile_w;
  c.h0 = th * P.tile_rows * P.mt;
  c.n0 = c.n_idx * P.n_tile;
  return c;
}

// v[0..31] += 32 consecutive bf16 channels at p (only the 8-channel groups below out_c)
__device__ __forceinline__ void add_tile32(float (&v)[32], const __nv_bfloat16* p, int ch, int out_c) {
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    if (ch + i * 8 < out_c) {
      const uint4 rv = __ldg(reinterpret_cast<const uint4*>(p + i * 8));
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

template <bool PAIR, bool AFFINE, bool SPLIT>
__global__ void __launch_bounds__(NTHREADS)
igemm_kernel(const __grid_constant__ IgemmMaps TM, const __grid_constant__ IgemmParams P) {
  const CUtensorMap& tmA0 = TM.a0[blockIdx.z];
  const CUtensorMap& tmA0h = TM.a0h[blockIdx.z];
  const CUtensorMap& tmA1 = TM.a1[blockIdx.z];
  const CUtensorMap& tmA1h = TM.a1h[blockIdx.z];
  const CUtensorMap& tmB = TM.b[blockIdx.z];
  const IgemmParams::Problem& Q = P.q[blockIdx.z];
  extern __shared__ __align__(1024) uint8_t smem[];   // swizzle atoms need 1024-byte alignment (checked below)
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)P.a_stages * P.a_slot_bytes;
  uint64_t* fullA = reinterpret_cast<uint64_t*>(sB + (size_t)P.b_stages * P.b_slot_bytes);
  uint64_t* emptyA = fullA + MAX_STAGES;
  uint64_t* fullB = emptyA + MAX_STAGES;
  uint64_t* emptyB = fullB + MAX_B;
  uint64_t* tmem_full = emptyB + MAX_B;    // [2]: accumulator buffer written, epilogue may drain it
  uint64_t* tmem_empty = tmem_full + 2;          // [2]: accumulator buffer drained, MMAs may overwrite it
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_empty + 2);
  float* s_stat = reinterpret_cast<float*>(tmem_holder + 4);   // [2][256] per-CTA channel sums

  const int warp = tc::warp_idx_uniform(), lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    trace_at(P, 0);
    if (tc::smem_u32(smem) & 1023u) {
      printf("margipose_b200: igemm shared memory base not 1024-byte aligned\n");
      __trap();
    }
  }
  if (Q.stat_sum)
    for (int i = threadIdx.x; i < 512; i += NTHREADS) s_stat[i] = 0.f;
  const uint32_t crank = PAIR ? tc::cluster_ctarank() : 0;
  // my contiguous range of work units (tiles, or tile pairs for a CTA pair)
  const int workers = PAIR ? (int)(gridDim.x >> 1) : (int)gridDim.x;
  const int me = PAIR ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  const int u_begin = (int)((long long)me * P.units / workers);
  const int u_end = (int)((long long)(me + 1) * P.units / workers);
  constexpr int STEP = PAIR ? 2 : 1;

  if (threadIdx.x == 0) {
    for (int s = 0; s < MAX_STAGES; ++s) {
      tc::mbar_init(&fullA[s], 1);
      tc::mbar_init(&emptyA[s], 1);
    }
    for (int s = 0; s < MAX_B; ++s) {
      tc::mbar_init(&fullB[s], 1);
      tc::mbar_init(&emptyB[s], 1);
    }
    for (int b = 0; b < 2; ++b) {
      tc::mbar_init(&tmem_full[b], 1);
      tc::mbar_init(&tmem_empty[b], PAIR ? 16 : 8);   // one arrival per epilogue warp (of both CTAs of a pair)
    }
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
  pdl_trigger();
  tc::tc_fence_before();
  __syncthreads();
  if (PAIR) tc::cluster_sync_all();   // the peer's barriers are initialised before anyone signals them
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_holder;
  // The packed weights are not produced by the stream predecessor (mp_pack_weights ran long before), so the first
  // weight tiles of this CTA's first work unit may be requested BEFORE the dependency wait: they land while the
  // previous kernel drains.  pre_b = weight-tile loads already issued, in the producer loop's (group, block, tap) order.
  int pre_b = 0;
  if (warp == 0 && u_begin < u_end && P.prefetch_b && !(P.dbg & 1)) {
    const TileCoord T = decode_tile(P, u_begin * STEP + (int)crank);
    const int b_row0 = T.n0 + (int)crank * (P.b_slot_bytes >> 7);
    for (int g = 0; g < P.n_groups && pre_b < P.b_stages; ++g) {
      const TapGroup& G = P.groups[g];
      for (int cb = 0; cb < P.cblocks && pre_b < P.b_stages; ++cb) {
        for (int i = 0; i < G.n && pre_b < P.b_stages; ++i) {
          uint8_t* dstB = sB + (size_t)pre_b * P.b_slot_bytes;
          if (tc::elect_one()) {
            if (PAIR) {
              if (crank == 0) tc::mbar_arrive_expect_tx(&fullB[pre_b], 2u * (uint32_t)P.b_slot_bytes);
              tc::tma_load_2d_pair(&tmB, &fullB[pre_b], dstB, G.koff[i] + cb * 64, b_row0);
            } else {
              tc::mbar_arrive_expect_tx(&fullB[pre_b], (uint32_t)P.b_slot_bytes);
              tc::tma_load_2d(&tmB, &fullB[pre_b], dstB, G.koff[i] + cb * 64, b_row0);
            }
          }
          ++pre_b;
        }
      }
    }
  }
  pdl_wait();   // everything above overlapped the previous kernel's tail; its results are visible from here on
  if (threadIdx.x == 0) trace_at(P, 1);

  if (warp == 0) {
    {   // ------------------------------------------- TMA producer: warp-uniform loops, one elected lane issues
      int sa = 0, sb = 0, b_issued = 0;
      uint32_t pha = 0, phb = 0;
      for (int u = u_begin; u < u_end; ++u) {
        const TileCoord T = decode_tile(P, u * STEP + (int)crank);
        if (P.resident) { sb = 0; phb = 0; }
        const bool load_b = !P.resident || u == u_begin;
        const int b_row0 = T.n0 + (int)crank * (P.b_slot_bytes >> 7);   // pair: my half of the weight-tile rows
        for (int g = 0; g < P.n_groups; ++g) {
          const TapGroup& G = P.groups[g];
          const bool halo = G.n > 1;
          const CUtensorMap* tmA = halo ? (G.src ? &tmA1h : &tmA0h) : (G.src ? &tmA1 : &tmA0);
          for (int cb = 0; cb < P.cblocks; ++cb) {
            tc::mbar_wait(&emptyA[sa], pha ^ 1);
            uint8_t* dstA = sA + (size_t)sa * P.a_slot_bytes;
            if (!tc::elect_one()) {
            } else if (P.dbg & 1) {
              if (crank == 0) tc::mbar_arrive(&fullA[sa]);
            } else if (PAIR) {   // both CTAs' bytes complete on the leader's barrier
              if (crank == 0) tc::mbar_arrive_expect_tx(&fullA[sa], 2u * (uint32_t)(halo ? P.a_tx_halo : P.a_tx_plain));
              tc::tma_load_5d_pair(tmA, &fullA[sa], dstA, G.c0 + cb * 64, T.w0 + G.dw, G.p, T.h0 + G.dh0, T.img);
            } else {
              tc::mbar_arrive_expect_tx(&fullA[sa], (uint32_t)(halo ? P.a_tx_halo : P.a_tx_plain));
              tc::tma_load_5d(tmA, &fullA[sa], dstA, G.c0 + cb * 64, T.w0 + G.dw, G.p, T.h0 + G.dh0, T.img);
            }
            if (++sa == P.a_stages) { sa = 0; pha ^= 1; }
            for (int i = 0; i < G.n && load_b; ++i) {
              if (b_issued < pre_b) {   // requested before the dependency wait (first ring pass: the slot was free)
                ++b_issued;
                if (++sb == P.b_stages) { sb = 0; phb ^= 1; }
                continue;
              }
              if (!P.resident) tc::mbar_wait(&emptyB[sb], phb ^ 1);
              uint8_t* dstB = sB + (size_t)sb * P.b_slot_bytes;
              if (!tc::elect_one()) {
              } else if (P.dbg & 1) {
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
    }
  } else if (warp == 1) {
    if (crank == 0) {   // ---------- MMA issuer (pair: the leader only): warp-uniform loops, one elected lane issues
      const uint32_t idesc = tc::idesc_bf16(PAIR ? 256 : 128, P.n_tile, false, false);
      int sa = 0, sb = 0;
      uint32_t pha = 0, phb = 0;
      int it = 0;
      for (int u = u_begin; u < u_end; ++u, ++it) {
        const int buf = it & 1;
        tc::mbar_wait(&tmem_empty[buf], (uint32_t)(((it >> 1) & 1) ^ 1));   // epilogue drained this buffer
        tc::tc_fence_after();
        const uint32_t acc = tmem + (uint32_t)(buf * P.acc_cols);
        bool first = true;
        if (P.resident) { sb = 0; phb = 0; }
        const bool wait_b = !P.resident || it == 0;
        for (int g = 0; g < P.n_groups; ++g) {
          const int n = P.groups[g].n;
          for (int cb = 0; cb < P.cblocks; ++cb) {
            tc::mbar_wait(&fullA[sa], pha);
            if (first && lane == 0) trace_at(P, 4 + it * 4);       // first operands of tile `it` have landed
            const uint32_t a0 = tc::smem_u32(sA + (size_t)sa * P.a_slot_bytes);
            for (int i = 0; i < n; ++i) {
              if (wait_b) tc::mbar_wait(&fullB[sb], phb);
              tc::tc_fence_after();
              const uint64_t bd = tc::desc_kmajor_sw128(tc::smem_u32(sB + (size_t)sb * P.b_slot_bytes));
              if (tc::elect_one()) {
                // row-shifted windows of the (halo) box: tap i of the group, row blocks 0 / 1 of the tile.  With two
                // row blocks the MMAs alternate between the two accumulators, so an MMA never waits for the one
                // just issued to the same accumulator.
                const uint64_t ad0 = tc::desc_kmajor_sw128(a0 + (uint32_t)(i * P.row_bytes));
                const uint64_t ad1 = tc::desc_kmajor_sw128(a0 + (uint32_t)((i + P.tile_rows) * P.row_bytes));
                const uint32_t d1 = acc + (uint32_t)P.n_tile;
#pragma unroll
                for (int k = 0; k < 4; ++k) {   // 4 x (K = 16 bf16 = 32 bytes inside the 128B swizzle atom: +2 in
                                                // the descriptor's 16-byte address field)
                  if (P.dbg & 2) continue;
                  if (PAIR) tc::mma_bf16_pair(acc, ad0 + 2 * k, bd + 2 * k, idesc, !(first && k == 0));
                  else tc::mma_bf16(acc, ad0 + 2 * k, bd + 2 * k, idesc, !(first && k == 0));
                  if (P.mt > 1) {
                    if (PAIR) tc::mma_bf16_pair(d1, ad1 + 2 * k, bd + 2 * k, idesc, !(first && k == 0));
                    else tc::mma_bf16(d1, ad1 + 2 * k, bd + 2 * k, idesc, !(first && k == 0));
                  }
                }
                // frees the weight slot (in both CTAs of a pair) once these MMAs have read it; with the last
                // tap of the group the activation slot too
                if (P.resident) {} else if (PAIR) tc::mma_commit_pair(&emptyB[sb]); else tc::mma_commit(&emptyB[sb]);
                if (i == n - 1) {
                  if (PAIR) tc::mma_commit_pair(&emptyA[sa]); else tc::mma_commit(&emptyA[sa]);
                }
              }
              first = false;
              if (++sb == P.b_stages) { sb = 0; phb ^= 1; }
            }
            if (++sa == P.a_stages) { sa = 0; pha ^= 1; }
          }
        }
        if (tc::elect_one()) {
          if (PAIR) tc::mma_commit_pair(&tmem_full[buf]); else tc::mma_commit(&tmem_full[buf]);
        }
        if (lane == 0) trace_at(P, 5 + it * 4);                    // all MMAs of tile `it` issued
      }
    }
  } else {   // ------------------------------------------------------------------------ epilogue
    const int q = warp & 3;
    const int et = threadIdx.x - 64;   // 0..255 among the epilogue threads
    const int half = (warp - 2) >> 2;  // which of the two warps of this TMEM lane quarter
    const int m = q * 32 + lane;
    const int r = m / P.tile_w, wq = m - r * P.tile_w;
    const int nchunks = P.n_tile / 32;
    int it = 0;
    for (int u = u_begin; u < u_end; ++u, ++it) {
      const TileCoord T = decode_tile(P, u * STEP + (int)crank);
      const int buf = it & 1;
      tc::mbar_wait(&tmem_full[buf], (uint32_t)((it >> 1) & 1));
      tc::tc_fence_after();
      if (et == 0) trace_at(P, 6 + it * 4);                        // accumulator of tile `it` complete
      const uint32_t acc = tmem + (uint32_t)(buf * P.acc_cols);
      const int n0 = T.n0, img = T.img;
      // The two warps of a TMEM lane quarter split the 32-column chunks by parity; a warp drains its chunk for
      // BOTH row blocks of the tile, so the BatchNorm sums of the two blocks are added per thread before the
      // (expensive) cross-lane transpose-reduction.
      for (int c = half; c < ((P.dbg & 4) ? 0 : nchunks); c += 2) {
        const int ch = n0 + c * 32;
        if (ch >= P.out_c) break;   // warp-uniform
        float s1[32], s2[32];
        if (Q.stat_sum) {
#pragma unroll
          for (int j = 0; j < 32; ++j) { s1[j] = 0.f; s2[j] = 0.f; }
        }
        for (int mt = 0; mt < P.mt; ++mt) {
          const int h = T.h0 + mt * P.tile_rows + r, w = T.w0 + wq;
          const bool valid = (r < P.tile_rows) && (h < P.out_h) && (w < P.out_w);
          const long long pix = (long long)img * P.out_sn + (long long)h * P.out_sh + (long long)w * P.out_sw;
          float v[32];
          tc::tmem_ld32(acc + ((uint32_t)(q * 32) << 16) + (uint32_t)(mt * P.n_tile + c * 32), v);
          if (!valid) continue;
          if (SPLIT && Q.acc_in) {   // the products of the other operand halves, accumulated by the earlier launches
            add_tile32(v, Q.acc_in + pix + ch, ch, P.out_c);
            add_tile32(v, Q.acc_in + pix + ch + P.lo_delta, ch, P.out_c);
          }
          if (AFFINE) {   // y = relu?(acc * scale[c] + shift[c]): warp-uniform addresses, 128-bit broadcast loads
#pragma unroll
            for (int i = 0; i < 8; ++i) {
              const float4 sc = __ldg(reinterpret_cast<const float4*>(Q.ep_scale + ch) + i);
              const float4 sh = __ldg(reinterpret_cast<const float4*>(Q.ep_shift + ch) + i);
              v[i * 4 + 0] = fmaf(v[i * 4 + 0], sc.x, sh.x);
              v[i * 4 + 1] = fmaf(v[i * 4 + 1], sc.y, sh.y);
              v[i * 4 + 2] = fmaf(v[i * 4 + 2], sc.z, sh.z);
              v[i * 4 + 3] = fmaf(v[i * 4 + 3], sc.w, sh.w);
            }
            if (P.ep_relu == 1) {
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
            }
          }
          if (Q.res) {
            add_tile32(v, Q.res + pix + ch, ch, P.out_c);
            if (SPLIT) add_tile32(v, Q.res + pix + ch + P.lo_delta, ch, P.out_c);
          }
          if (AFFINE && P.ep_relu == 2) {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = fmaxf(v[j], 0.f);
          }
          uint32_t packed[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) packed[j] = pack_bf16x2(v[2 * j], v[2 * j + 1]);
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (ch + i * 8 < P.out_c)
              *reinterpret_cast<uint4*>(Q.out + pix + ch + i * 8) =
                  make_uint4(packed[i * 4], packed[i * 4 + 1], packed[i * 4 + 2], packed[i * 4 + 3]);
          }
          if (SPLIT) {   // second bf16 of the pair: what the first rounding lost (value = hi + lo, ~16 significant bits)
            uint32_t plo[16];
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 f = unpack_bf16x2(packed[j]);
              plo[j] = pack_bf16x2(v[2 * j] - f.x, v[2 * j + 1] - f.y);
              const float2 g = unpack_bf16x2(plo[j]);
              v[2 * j] = f.x + g.x;           // the value as stored, for the statistics below
              v[2 * j + 1] = f.y + g.y;
            }
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              if (ch + i * 8 < P.out_c)
                *reinterpret_cast<uint4*>(Q.out + pix + ch + i * 8 + P.lo_delta) =
                    make_uint4(plo[i * 4], plo[i * 4 + 1], plo[i * 4 + 2], plo[i * 4 + 3]);
            }
          }
          if (Q.stat_sum) {   // statistics of the values as stored (bf16-rounded; split mode: hi + lo)
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              float2 f = unpack_bf16x2(packed[j]);
              if (SPLIT) f = make_float2(v[2 * j], v[2 * j + 1]);
              s1[2 * j] += f.x; s1[2 * j + 1] += f.y;
              s2[2 * j] = fmaf(f.x, f.x, s2[2 * j]); s2[2 * j + 1] = fmaf(f.y, f.y, s2[2 * j + 1]);
            }
          }
        }
        if (Q.stat_sum) {
          const float a1 = tc::warp_transpose_sum(s1);
          const float a2 = tc::warp_transpose_sum(s2);
          atomicAdd(s_stat + c * 32 + lane, a1);          // the epilogue warps meet in shared memory ...
          atomicAdd(s_stat + 256 + c * 32 + lane, a2);
        }
      }
      // this warp's share of the buffer is in registers / stored: hand it back to the MMA issuer
      tc::tc_fence_before();
      __syncwarp();
      if (lane == 0) {
        if (PAIR) tc::mbar_arrive_remote(&tmem_empty[buf], 0); else tc::mbar_arrive(&tmem_empty[buf]);
      }
      if (et == 0) trace_at(P, 7 + it * 4);                        // tile `it` drained (this warp)
      // BatchNorm sums leave the CTA when its range ends or moves on to other output channels
      const bool flush = (u + 1 == u_end) || ((u + 1) * STEP / P.m_tiles != T.n_idx);
      if (Q.stat_sum && flush) {
        asm volatile("bar.sync 1, 256;" ::: "memory");   // the 4 epilogue warps have met in shared memory
        const long long rep = (long long)(blockIdx.x % P.stat_replicas) * P.stat_stride;
        for (int i = et; i < P.n_tile; i += EPI_THREADS) {
          if (n0 + i < P.out_c) {   // ONE global atomic per channel and statistic
            atomicAdd(Q.stat_sum + rep + n0 + i, s_stat[i]);
            atomicAdd(Q.stat_sq + rep + n0 + i, s_stat[256 + i]);
          }
        }
        if (Q.fin_counter) {   // last CTA to arrive turns the sums into the BatchNorm coefficients
          __threadfence();
          asm volatile("bar.sync 1, 256;" ::: "memory");
          unsigned* s_ticket = reinterpret_cast<unsigned*>(s_stat + 512);
          if (et == 0) *s_ticket = atomicAdd(Q.fin_counter, 1u);
          asm volatile("bar.sync 1, 256;" ::: "memory");
          if (*s_ticket == (unsigned)(P.fin_total - 1)) {
            __threadfence();
          const float inv_m = 1.0f / (float)P.fin_count;
          for (int c = et; c < P.fin_Cp; c += EPI_THREADS) {
            float scale = 0.f, shift = 0.f;
            if (c < P.fin_C) {
              // every load before the first store: one L2 / DRAM round trip on the launch's serial tail instead of two
              const float s1 = __ldcg(Q.stat_sum + c), s2 = __ldcg(Q.stat_sq + c);
              const float gamma = Q.fin_gamma[c], beta = Q.fin_beta[c];
              const float rm = Q.fin_rmean ? Q.fin_rmean[c] : 0.f, rv = Q.fin_rmean ? Q.fin_rvar[c] : 0.f;
              const float bias = Q.fin_bias ? Q.fin_bias[c] : 0.f;
              const float mean = s1 * inv_m;
              const float var = fmaxf(s2 * inv_m - mean * mean, 0.f);
              const float invstd = rsqrtf(var + P.fin_eps);
              scale = gamma * invstd;
              shift = fmaf(-mean, scale, beta);
              Q.fin_smean[c] = mean;
              Q.fin_sinvstd[c] = invstd;
              if (Q.fin_rmean) {
                const float unbiased = P.fin_count > 1 ? var * ((float)P.fin_count / (float)(P.fin_count - 1)) : var;
                Q.fin_rmean[c] = (1.f - P.fin_momentum) * rm + P.fin_momentum * (mean + bias);
                Q.fin_rvar[c] = (1.f - P.fin_momentum) * rv + P.fin_momentum * unbiased;
              }
            }
            Q.fin_scale[c] = scale;
            Q.fin_shift[c] = shift;
          }
          }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        for (int i = et; i < 512; i += EPI_THREADS) s_stat[i] = 0.f;
        asm volatile("bar.sync 1, 256;" ::: "memory");
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (threadIdx.x == 0) trace_at(P, 2);
  if (PAIR) {
    tc::cluster_sync_all();   // nobody leaves while the pair's MMAs may still read its shared memory / TMEM
    if (warp == 1) tc::tmem_dealloc_pair(tmem, P.tmem_cols);
  } else if (warp == 1) {
    tc::tmem_dealloc(tmem, P.tmem_cols);
  }
}

long long g_igemm_smem = 227 * 1024;   // per CTA (one persistent CTA per SM): activation ring + weight ring / resident weights
long long g_igemm_astages = 3;     // activation (halo box) slots when the weights stream
long long g_igemm_resident = 1;    // keep the weight operand in shared memory across a CTA's tiles when it fits
long long g_igemm_ctas = 0;        // CTAs per launch (0 = the device's SM count)
long long g_igemm_halo = 1;        // group taps that differ only in their row shift (one halo box per group)
long long g_igemm_dbg = 0;
// first weight tiles requested before the programmatic-dependency wait: measured 2 % SLOWER in the step (serial igemm
// 5.39 -> 5.50 ms; the early loads compete with the predecessor's tail and lengthen the prologue), hence off
long long g_igemm_prefetch_b = 0;
long long g_igemm_trace = 0;
long long g_igemm_pair = 1;        // CTA pairs (cta_group::2) when the M tiles pair up
long long g_igemm_split_n = 0;     // split N in two when a launch has fewer tiles than this (0 = the CTA count)
long long g_igemm_mt = 2;          // row blocks (accumulators) per tile when the launch keeps >= g_igemm_mt_ctas tiles
long long g_igemm_mt_ctas = 0;     // (0 = the CTA count)
bool g_attr_set = false;
int g_sm_count = 0;

int cta_budget() {
  if (g_igemm_ctas > 0) return (int)g_igemm_ctas;
  if (g_sm_count == 0) {
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess &&
        n > 0)
      g_sm_count = n;
    else
      g_sm_count = 148;
  }
  return g_sm_count;
}

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

// Partition the taps into groups of up to three that share (src, c0, dw, p) and have consecutive row shifts (taken
// in list order, so the hi*hi / lo*hi / hi*lo blocks of a split-mode tap list group among themselves); everything else
// becomes a group of one.  Returns the number of groups.
int group_taps(const mp_igemm_args* a, bool allow_halo, TapGroup* out) {
  bool used[MP_MAX_TAPS] = {};
  int n_groups = 0;
  for (int i = 0; i < a->n_taps; ++i) {
    if (used[i]) continue;
    const mp_tap& t = a->taps[i];
    used[i] = true;
    int dh[3] = {t.dh, 0, 0}, koff[3] = {t.koff, 0, 0}, n = 1;
    int lo = t.dh, hi = t.dh;
    if (allow_halo) {
      bool grew = true;
      while (grew && n < 3) {
        grew = false;
        for (int j = 0; j < a->n_taps && n < 3; ++j) {
          const mp_tap& u = a->taps[j];
          if (used[j] || u.src != t.src || u.c0 != t.c0 || u.dw != t.dw || u.p != t.p) continue;
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
void mp_set_igemm_prefetch_b(long long v) { g_igemm_prefetch_b = v; }
void mp_set_igemm_trace(long long v) { g_igemm_trace = v; }
void mp_set_igemm_resident(long long v) { g_igemm_resident = v; }
void mp_set_igemm_astages(long long v) { g_igemm_astages = v < 2 ? 2 : v; }
void mp_set_igemm_split_n(long long v) { g_igemm_split_n = v; }
void mp_set_igemm_mt(int which, long long v) {
  if (which == 0) g_igemm_mt = v;
  else if (which == 1) g_igemm_mt_ctas = v;
  else g_igemm_ctas = v;
}

static void igemm_grid(const mp_igemm_args* a, int n_problems, int* tile_w, int* tile_rows, int* tiles_w, int* tiles_h,
                       int* n_tile, int* mt) {
  mp_pick_tile(a->out_h, a->out_w, 128, tile_w, tile_rows);
  *tiles_w = (a->out_w + *tile_w - 1) / *tile_w;
  const int budget = cta_budget();
  // Preferred shape: TWO 128-pixel row blocks per tile, i.e. two TMEM accumulators that share every weight
  // tile and that the MMAs alternate between (back-to-back MMAs into ONE accumulator run at about half the
  // tensor rate when N <= 128).  Two accumulators, double-buffered, need 4 * n_tile <= 512 TMEM columns,
  // so wider outputs are split along N (the activations are then streamed once per part).
  *mt = 1;
  *n_tile = a->w_rows <= 256 ? (int)a->w_rows : 256;
  if (g_igemm_mt >= 2 && *tile_w * *tile_rows == 128 && a->out_h >= 2 * *tile_rows) {
    int nt = *n_tile;
    while (nt > 128 && nt % 2 == 0 && (nt / 2) % 32 == 0) nt /= 2;
    const long long th2 = (a->out_h + 2 * *tile_rows - 1) / (2 * *tile_rows);
    const long long need = g_igemm_mt_ctas > 0 ? g_igemm_mt_ctas : budget / 2;
    if (nt <= 128 && a->w_rows % nt == 0 &&
        (long long)n_problems * a->n_img * th2 * *tiles_w * (a->w_rows / nt) >= need) {
      *mt = 2;
      *n_tile = nt;
    }
  }
  *tiles_h = (a->out_h + *tile_rows * *mt - 1) / (*tile_rows * *mt);
  // Small pixel grids leave SMs idle or badly balanced: give each pixel tile two work units with half the
  // output channels each.
  const long long tiles = (long long)n_problems * a->n_img * *tiles_h * *tiles_w * (a->w_rows / *n_tile);
  const long long want = g_igemm_split_n > 0 ? g_igemm_split_n : budget;
  if (*mt == 1 && tiles < want && *n_tile >= 128 && (*n_tile / 2) % 32 == 0) *n_tile /= 2;
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
      (x->ep_scale == nullptr) != (y->ep_scale == nullptr) || x->ep_relu != y->ep_relu ||
      (x->acc_in == nullptr) != (y->acc_in == nullptr) || x->lo_delta != y->lo_delta ||
      x->stat_stride != y->stat_stride || !same_view(x->src[0], y->src[0]) ||
      (x->src[1].ptr == nullptr) != (y->src[1].ptr == nullptr) || (x->src[1].ptr && !same_view(x->src[1], y->src[1])))
    return false;
  if (x->bn && (x->bn_launches != y->bn_launches || x->bn_channels != y->bn_channels || x->bn_count != y->bn_count ||
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
  igemm_grid(a, n_problems, &P.tile_w, &P.tile_rows, &P.tiles_w, &P.tiles_h, &P.n_tile, &P.mt);
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

  // CTA pairs: adjacent pixel tiles, each CTA holds half of the weight-tile rows
  P.m_tiles = a->n_img * P.tiles_h * P.tiles_w;
  auto pair_ok = [&](int n_tile) {
    return g_igemm_pair != 0 && P.m_tiles % 2 == 0 && n_tile % 16 == 0 && (n_tile / 2) % 8 == 0;
  };
  // Resident weights: when this CTA's share of the whole weight operand fits beside two activation slots it
  // is loaded once (with the first tile) and later tiles stream activations only.
  const int overhead = 3072;   // barriers + TMEM address + per-CTA channel sums + ticket
  const long long budget = g_igemm_smem - overhead;
  const int b_slots = a->n_taps * a->cblocks;
  auto fits = [&](int n_tile) {
    const long long bytes = (long long)b_slots * (n_tile / (pair_ok(n_tile) ? 2 : 1)) * 128;
    return b_slots <= MAX_B && 2LL * P.a_slot_bytes + bytes <= budget;
  };
  P.resident = 0;
  if (g_igemm_resident) {
    if (fits(P.n_tile)) P.resident = 1;
  }
  const int pair = pair_ok(P.n_tile) ? 1 : 0;
  const int n_tiles = (int)(a->w_rows / P.n_tile);
  P.acc_cols = P.mt * P.n_tile;
  {
    const int cols = 2 * P.acc_cols;   // double-buffered accumulator
    MP_CHECK_ARG(cols <= 512, "mp_conv_igemm: %d TMEM columns needed", cols);
    P.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
  }

  // work units and CTAs: every problem gets the same number of persistent CTAs, each a contiguous range
  P.units = P.m_tiles * n_tiles / (pair ? 2 : 1);
  int workers = cta_budget() / n_problems / (pair ? 2 : 1);
  if (workers < 1) workers = 1;
  if (workers > P.units) workers = P.units;
  if (P.resident && n_tiles > 1) {   // a CTA's range must stay within one output-channel block
    if (workers >= n_tiles) workers -= workers % n_tiles;
    else P.resident = 0;
  }
  P.b_slot_bytes = (P.n_tile / (pair ? 2 : 1)) * 128;
  P.dbg = (int)g_igemm_dbg;
  P.prefetch_b = (int)g_igemm_prefetch_b;
  P.trace = reinterpret_cast<unsigned long long*>(g_igemm_trace);

  // shared-memory rings: with halo groups each activation box feeds up to three weight tiles
  int a_stages, b_stages;
  if (P.resident) {
    b_stages = b_slots;
    a_stages = (int)((budget - (long long)b_slots * P.b_slot_bytes) / P.a_slot_bytes);
  } else if (any_halo) {
    a_stages = (int)g_igemm_astages;   // activation boxes in flight; the weight ring gets the rest
    while (a_stages > 2 && budget - (long long)a_stages * P.a_slot_bytes < 4LL * P.b_slot_bytes) --a_stages;
    b_stages = (int)((budget - (long long)a_stages * P.a_slot_bytes) / P.b_slot_bytes);
    if (b_stages > MAX_STAGES) b_stages = MAX_STAGES;
  } else {
    a_stages = b_stages = (int)(budget / (P.a_slot_bytes + P.b_slot_bytes));
    if (b_stages > MAX_STAGES) b_stages = MAX_STAGES;
  }
  if (a_stages < 2) a_stages = 2;
  if (b_stages < 2) b_stages = 2;
  if (a_stages > MAX_STAGES) a_stages = MAX_STAGES;
  P.a_stages = a_stages;
  P.b_stages = b_stages;
  P.out_sn = a->out_sn; P.out_sh = a->out_sh; P.out_sw = a->out_sw;
  P.out_c = a->out_c;
  P.stat_replicas = a->stat_replicas > 1 ? a->stat_replicas : 1;
  P.stat_stride = a->stat_stride;
  P.ep_relu = a->ep_relu;
  P.lo_delta = a->lo_delta;
  MP_CHECK_ARG(a->lo_delta >= 0 && a->lo_delta % 8 == 0, "mp_conv_igemm: lo_delta must be a non-negative multiple of 8");
  MP_CHECK_ARG(a->lo_delta > 0 || a->acc_in == nullptr, "mp_conv_igemm: acc_in needs split mode (lo_delta > 0)");
  MP_CHECK_ARG(a->ep_relu >= 0 && a->ep_relu <= 2, "mp_conv_igemm: ep_relu %d out of range", a->ep_relu);
  P.fin_total = 0; P.fin_C = 0; P.fin_Cp = 0; P.fin_count = 0; P.fin_momentum = 0.f; P.fin_eps = 0.f;
  if (a->bn) {
    MP_CHECK_ARG(a->bn_count > 0 && a->bn_channels > 0 && a->bn_channels <= a->out_c,
                 "mp_conv_igemm: incomplete BatchNorm finalize arguments");
    P.fin_C = a->bn_channels; P.fin_Cp = a->out_c;
    P.fin_count = a->bn_count; P.fin_momentum = a->bn_momentum; P.fin_eps = a->bn_eps;
    // arrivals per problem and launch: one per CTA and run of tiles with the same output-channel block
    long long arrivals = 0;
    const int step = pair ? 2 : 1;
    for (int w = 0; w < workers; ++w) {
      const long long ub = (long long)w * P.units / workers, ue = (long long)(w + 1) * P.units / workers;
      if (ue > ub) arrivals += ((ue - 1) * step / P.m_tiles - ub * step / P.m_tiles + 1) * step;
    }
    P.fin_total = (int)(arrivals * (a->bn_launches > 1 ? a->bn_launches : 1));
  }

  IgemmMaps TM;
  const uint32_t boxA[5] = {64, (uint32_t)P.tile_w, 1, (uint32_t)(P.tile_rows * P.mt), 1};
  const uint32_t boxH[5] = {64, (uint32_t)P.tile_w, 1, (uint32_t)(P.tile_rows * P.mt + 2), 1};
  for (int i = 0; i < MP_MAX_GROUP; ++i) {
    const mp_igemm_args* x = &args[i < n_problems ? i : 0];
    IgemmParams::Problem& Q = P.q[i];
    Q.out = reinterpret_cast<__nv_bfloat16*>(x->out);
    Q.res = reinterpret_cast<const __nv_bfloat16*>(x->res);
    Q.acc_in = reinterpret_cast<const __nv_bfloat16*>(x->acc_in);
    Q.stat_sum = x->stat_sum;
    Q.stat_sq = x->stat_sq;
    Q.fin_gamma = Q.fin_beta = Q.fin_bias = nullptr;
    Q.fin_rmean = Q.fin_rvar = Q.fin_smean = Q.fin_sinvstd = Q.fin_scale = Q.fin_shift = nullptr;
    Q.fin_counter = nullptr;
    Q.ep_scale = x->ep_scale;
    Q.ep_shift = x->ep_shift;
    MP_CHECK_ARG((x->ep_scale == nullptr) == (x->ep_shift == nullptr) &&
                     (!x->ep_scale || (mp_aligned16(x->ep_scale) && mp_aligned16(x->ep_shift))),
                 "mp_conv_igemm: ep_scale / ep_shift go together and must be 16-byte aligned");
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
      if (any_halo) {
        rc = view_to_tmap(&TM.a1h[i], x->src[1], boxH, "mp_conv_igemm src[1] (halo box)");
        if (rc != MP_OK) return rc;
      } else {
        TM.a1h[i] = TM.a1[i];
      }
    } else {
      TM.a1[i] = TM.a0[i];
      TM.a1h[i] = TM.a0h[i];
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
#define MP_IGEMM_ATTR(P_, A_, S_) \
    MP_CUDA(cudaFuncSetAttribute(igemm_kernel<P_, A_, S_>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024))
    MP_IGEMM_ATTR(false, false, false); MP_IGEMM_ATTR(true, false, false);
    MP_IGEMM_ATTR(false, true, false); MP_IGEMM_ATTR(true, true, false);
    MP_IGEMM_ATTR(false, false, true); MP_IGEMM_ATTR(true, false, true);
    MP_IGEMM_ATTR(false, true, true); MP_IGEMM_ATTR(true, true, true);
#undef MP_IGEMM_ATTR
    g_attr_set = true;
  }
  dim3 grid((unsigned)(workers * (pair ? 2 : 1)), 1, (unsigned)n_problems);
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = pair ? 2 : 1;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = mp_pdl_enabled() ? 2 : 1;
  const bool affine = a->ep_scale != nullptr;
  MP_CHECK_ARG(affine || a->ep_relu == 0, "mp_conv_igemm: ep_relu needs ep_scale / ep_shift");
  const bool split = a->lo_delta > 0;
#define MP_IGEMM_GO(P_, A_, S_) MP_CUDA(cudaLaunchKernelEx(&cfg, igemm_kernel<P_, A_, S_>, TM, P))
  if (split) {
    if (pair && affine) MP_IGEMM_GO(true, true, true);
    else if (pair) MP_IGEMM_GO(true, false, true);
    else if (affine) MP_IGEMM_GO(false, true, true);
    else MP_IGEMM_GO(false, false, true);
  } else {
    if (pair && affine) MP_IGEMM_GO(true, true, false);
    else if (pair) MP_IGEMM_GO(true, false, false);
    else if (affine) MP_IGEMM_GO(false, true, false);
    else MP_IGEMM_GO(false, false, false);
  }
#undef MP_IGEMM_GO
  MP_CHECK_LAUNCH("mp_conv_igemm");
  return MP_OK;
}
