// Convolution forward / data-gradient as an implicit GEMM on tcgen05 (sm_100a).
//
//   D[pixel, n] = sum_taps sum_c  A_tap[pixel, c] * W[n, koff_tap + c]
//
// A_tap is never materialised: one TMA box {64 ch, tile_w, 1, tile_rows, 1} of the 5-D activation
// view lands the tap's shifted (and, for stride 2, decimated) 128-pixel x 64-channel operand tile
// in shared memory in the 128B-swizzled K-major layout tcgen05.mma reads; image borders are TMA
// out-of-bounds zero fill.  W tiles ([n_tile rows] x 64 K) arrive the same way.  One elected
// thread issues tcgen05.mma (M=128, N=n_tile, K=16) into a TMEM accumulator; four epilogue warps
// pull it back with tcgen05.ld, add the optional residual, round to bf16, store NHWC and reduce
// the BatchNorm batch statistics of the stored tile.
//
// Replaces the cuDNN calls behind nn.Conv2d / nn.ConvTranspose2d on the reference's hot path
// (/root/reference/src/margipose/models/margipose_model.py:33,67-68,73-74,79-82 and the
// torchvision ResNet stem :130-135), forward and dgrad.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = TMEM owner + MMA issuer,
// warps 2-5 = epilogue (TMEM lane quarter = warp % 4).  Roofline: tensor pipe; algorithmic
// flops per launch = 2 * pixels * n * K.
#include "tc.cuh"
#include "../../include/margipose_b200.h"

namespace {

constexpr int NTHREADS = 192;
constexpr int A_STAGE_BYTES = 128 * 128;   // 128 pixels x 64 bf16

struct IgemmParams {
  mp_tap taps[MP_MAX_TAPS];
  int n_taps, cblocks;
  int tile_w, tile_rows, tiles_w, tiles_h;
  int out_h, out_w;
  int n_tile, stages, b_stage_bytes, stage_tx, tmem_cols;
  int cluster;   // CTAs per cluster along the M tiles (1, 2 or 4): each loads 1/cluster of the weight tile, multicast to all
  __nv_bfloat16* out;
  const __nv_bfloat16* res;
  long long out_sn, out_sh, out_sw;
  int out_c;
  float* stat_sum;
  float* stat_sq;
  int stat_replicas;
  long long stat_stride;
  // BatchNorm finalize by the last CTA (see mp_igemm_args.bn)
  const float* fin_gamma; const float* fin_beta; const float* fin_bias;
  float* fin_rmean; float* fin_rvar; float* fin_smean; float* fin_sinvstd; float* fin_scale; float* fin_shift;
  unsigned* fin_counter;
  int fin_total, fin_C, fin_Cp;
  long long fin_count;
  float fin_momentum, fin_eps;
};

__global__ void __launch_bounds__(NTHREADS)
igemm_kernel(const __grid_constant__ CUtensorMap tmA0, const __grid_constant__ CUtensorMap tmA1,
             const __grid_constant__ CUtensorMap tmB, const __grid_constant__ IgemmParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = P.stages;
  uint8_t* sA = smem;
  uint8_t* sB = smem + (size_t)stages * A_STAGE_BYTES;
  uint64_t* full = reinterpret_cast<uint64_t*>(sB + (size_t)stages * P.b_stage_bytes);
  uint64_t* empty = full + stages;
  uint64_t* tmem_full = empty + stages;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full + 1);
  float* s_stat = reinterpret_cast<float*>(tmem_holder + 4);   // [2][256] per-CTA channel sums

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (P.stat_sum)
    for (int i = threadIdx.x; i < 512; i += NTHREADS) s_stat[i] = 0.f;
  int t = blockIdx.x;
  const int tw = t % P.tiles_w; t /= P.tiles_w;
  const int th = t % P.tiles_h;
  const int img = t / P.tiles_h;
  const int w0 = tw * P.tile_w, h0 = th * P.tile_rows;
  const int n0 = blockIdx.y * P.n_tile;

  const int C = P.cluster;
  const uint32_t crank = C > 1 ? tc::cluster_ctarank() : 0;
  const uint16_t cmask = (uint16_t)((1u << C) - 1);
  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], C);   // every CTA of the cluster must have consumed a slot peers multicast into
    }
    tc::mbar_init(tmem_full, 1);
    tc::mbar_fence_init();
    tc::prefetch_tmap(&tmA0);
    tc::prefetch_tmap(&tmA1);
    tc::prefetch_tmap(&tmB);
  }
  if (warp == 1) tc::tmem_alloc(tmem_holder, P.tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  if (C > 1) tc::cluster_sync_all();   // peers' barriers are initialised before anyone signals them
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {   // ---------------------------------------------------------- TMA producer
      int s = 0;
      uint32_t ph = 0;
      for (int tp = 0; tp < P.n_taps; ++tp) {
        const mp_tap tap = P.taps[tp];
        const CUtensorMap* tmA = tap.src ? &tmA1 : &tmA0;
        for (int cb = 0; cb < P.cblocks; ++cb) {
          tc::mbar_wait(&empty[s], ph ^ 1);
          tc::mbar_arrive_expect_tx(&full[s], P.stage_tx);
          tc::tma_load_5d(tmA, &full[s], sA + (size_t)s * A_STAGE_BYTES, tap.c0 + cb * 64, w0 + tap.dw,
                          tap.p, h0 + tap.dh, img);
          if (C == 1) {
            tc::tma_load_2d(&tmB, &full[s], sB + (size_t)s * P.b_stage_bytes, tap.koff + cb * 64, n0);
          } else {   // my 1/C of the weight tile goes to every CTA of the cluster
            const int rows = P.n_tile / C;
            tc::tma_load_2d_multicast(&tmB, &full[s], sB + (size_t)s * P.b_stage_bytes + (size_t)crank * rows * 128,
                                      tap.koff + cb * 64, n0 + (int)crank * rows, cmask);
          }
          if (++s == stages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ------------------------------------------------------------ MMA issuer
      const uint32_t idesc = tc::idesc_bf16(128, P.n_tile, false, false);
      const int KB = P.n_taps * P.cblocks;
      int s = 0;
      uint32_t ph = 0;
      for (int kb = 0; kb < KB; ++kb) {
        tc::mbar_wait(&full[s], ph);
        tc::tc_fence_after();
        const uint32_t a = tc::smem_u32(sA + (size_t)s * A_STAGE_BYTES);
        const uint32_t b = tc::smem_u32(sB + (size_t)s * P.b_stage_bytes);
#pragma unroll
        for (int k = 0; k < 4; ++k)   // 4 x (K = 16 bf16 = 32 bytes inside the 128B swizzle atom)
          tc::mma_bf16(tmem, tc::desc_kmajor_sw128(a + k * 32), tc::desc_kmajor_sw128(b + k * 32), idesc,
                       (kb | k) != 0);
        if (C == 1) tc::mma_commit(&empty[s]);   // frees the smem slot once these MMAs have read it
        else tc::mma_commit_multicast(&empty[s], cmask);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
      tc::mma_commit(tmem_full);
    }
  } else {   // ------------------------------------------------------------------------ epilogue
    tc::mbar_wait(tmem_full, 0);
    tc::tc_fence_after();
    const int q = warp & 3;
    const int m = q * 32 + lane;
    const int r = m / P.tile_w, wq = m - r * P.tile_w;
    const int h = h0 + r, w = w0 + wq;
    const bool valid = (r < P.tile_rows) && (h < P.out_h) && (w < P.out_w);
    const long long pix = (long long)img * P.out_sn + (long long)h * P.out_sh + (long long)w * P.out_sw;
    const int nchunks = P.n_tile / 32;
    for (int c = 0; c < nchunks; ++c) {
      const int ch = n0 + c * 32;
      if (ch >= P.out_c) break;   // warp-uniform
      float v[32];
      tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(c * 32), v);
      uint32_t packed[16];
      if (valid) {
        if (P.res) {
#pragma unroll
          for (int i = 0; i < 4; ++i) {
            if (ch + i * 8 < P.out_c) {
              const uint4 rv = __ldg(reinterpret_cast<const uint4*>(P.res + pix + ch + i * 8));
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
            *reinterpret_cast<uint4*>(P.out + pix + ch + i * 8) =
                make_uint4(packed[i * 4], packed[i * 4 + 1], packed[i * 4 + 2], packed[i * 4 + 3]);
        }
      }
      if (P.stat_sum) {   // statistics of the values as stored (bf16-rounded)
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
    if (P.stat_sum) {   // ... and the CTA issues ONE global atomic per channel and statistic
      asm volatile("bar.sync 1, 128;" ::: "memory");
      const long long rep = (long long)(blockIdx.x % P.stat_replicas) * P.stat_stride;
      for (int i = threadIdx.x - 64; i < P.n_tile; i += 128) {
        if (n0 + i < P.out_c) {
          atomicAdd(P.stat_sum + rep + n0 + i, s_stat[i]);
          atomicAdd(P.stat_sq + rep + n0 + i, s_stat[256 + i]);
        }
      }
      if (P.fin_counter) {   // last CTA to arrive turns the sums into the BatchNorm coefficients
        __threadfence();
        asm volatile("bar.sync 1, 128;" ::: "memory");
        unsigned* s_ticket = reinterpret_cast<unsigned*>(s_stat + 512);
        if (threadIdx.x == 64) *s_ticket = atomicAdd(P.fin_counter, 1u);
        asm volatile("bar.sync 1, 128;" ::: "memory");
        if (*s_ticket == (unsigned)(P.fin_total - 1)) {
          __threadfence();
          const float inv_m = 1.0f / (float)P.fin_count;
          for (int c = threadIdx.x - 64; c < P.fin_Cp; c += 128) {
            float scale = 0.f, shift = 0.f;
            if (c < P.fin_C) {
              const float mean = __ldcg(P.stat_sum + c) * inv_m;
              const float var = fmaxf(__ldcg(P.stat_sq + c) * inv_m - mean * mean, 0.f);
              const float invstd = rsqrtf(var + P.fin_eps);
              scale = P.fin_gamma[c] * invstd;
              shift = fmaf(-mean, scale, P.fin_beta[c]);
              P.fin_smean[c] = mean;
              P.fin_sinvstd[c] = invstd;
              if (P.fin_rmean) {
                const float unbiased = P.fin_count > 1 ? var * ((float)P.fin_count / (float)(P.fin_count - 1)) : var;
                const float bias = P.fin_bias ? P.fin_bias[c] : 0.f;
                P.fin_rmean[c] = (1.f - P.fin_momentum) * P.fin_rmean[c] + P.fin_momentum * (mean + bias);
                P.fin_rvar[c] = (1.f - P.fin_momentum) * P.fin_rvar[c] + P.fin_momentum * unbiased;
              }
            }
            P.fin_scale[c] = scale;
            P.fin_shift[c] = shift;
          }
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (C > 1) tc::cluster_sync_all();   // nobody leaves while peers may still signal its barriers
  if (warp == 1) tc::tmem_dealloc(tmem, P.tmem_cols);
}

long long g_igemm_smem = 115712;   // 113 KB: two CTAs share an SM (one's epilogue overlaps the other's main loop)
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

long long g_igemm_cluster = 1;    // max CTAs per cluster sharing a multicast weight tile (measured: no gain, the
                                  // kernel is bound by shared-memory bandwidth, not by L2 -> SM traffic)
void mp_set_igemm_cluster(long long v) { g_igemm_cluster = v; }
long long g_igemm_split_n = 100;   // split N in two when the launch would have fewer CTAs than this
void mp_set_igemm_split_n(long long v) { g_igemm_split_n = v; }

static void igemm_grid(const mp_igemm_args* a, int* tile_w, int* tile_rows, int* tiles_w, int* tiles_h, int* n_tile) {
  mp_pick_tile(a->out_h, a->out_w, 128, tile_w, tile_rows);
  *tiles_w = (a->out_w + *tile_w - 1) / *tile_w;
  *tiles_h = (a->out_h + *tile_rows - 1) / *tile_rows;
  *n_tile = a->w_rows <= 256 ? (int)a->w_rows : 256;
  // Small pixel grids (e.g. 16x16 maps at batch 32 = 64 M-tiles) leave most of the 148 SMs idle:
  // give each M-tile two CTAs with half the output channels each (the A tile is then fetched twice,
  // from L2, which is cheaper than idle tensor cores).
  const long long ctas = (long long)a->n_img * *tiles_h * *tiles_w * (a->w_rows / *n_tile);
  if (ctas < g_igemm_split_n && *n_tile >= 128 && (*n_tile / 2) % 32 == 0) *n_tile /= 2;
}

extern "C" int mp_conv_igemm_ctas(const mp_igemm_args* a) {
  if (!a || a->out_h <= 0 || a->out_w <= 0 || a->w_rows <= 0) return 0;
  int tw, tr, tsw, tsh, nt;
  igemm_grid(a, &tw, &tr, &tsw, &tsh, &nt);
  return a->n_img * tsh * tsw * (int)(a->w_rows / nt);
}

extern "C" int mp_conv_igemm(const mp_igemm_args* a, void* stream) {
  MP_CHECK_ARG(a, "mp_conv_igemm: null args");
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
  bool use_src1 = false;
  for (int i = 0; i < a->n_taps; ++i) {
    MP_CHECK_ARG(a->taps[i].src == 0 || a->taps[i].src == 1, "mp_conv_igemm: tap %d: bad src", i);
    MP_CHECK_ARG(a->taps[i].koff >= 0 && a->taps[i].koff + a->cblocks * 64 <= a->w_k,
                 "mp_conv_igemm: tap %d: K range outside the weight matrix", i);
    if (a->taps[i].src == 1) use_src1 = true;
  }
  MP_CHECK_ARG(!use_src1 || a->src[1].ptr, "mp_conv_igemm: tap refers to a missing second source");

  IgemmParams P;
  for (int i = 0; i < a->n_taps; ++i) P.taps[i] = a->taps[i];
  P.n_taps = a->n_taps;
  P.cblocks = a->cblocks;
  igemm_grid(a, &P.tile_w, &P.tile_rows, &P.tiles_w, &P.tiles_h, &P.n_tile);
  P.out_h = a->out_h;
  P.out_w = a->out_w;
  MP_CHECK_ARG(a->w_rows % P.n_tile == 0, "mp_conv_igemm: w_rows %lld not tileable", (long long)a->w_rows);
  P.b_stage_bytes = P.n_tile * 128;
  const int stage_bytes = A_STAGE_BYTES + P.b_stage_bytes;
  const int overhead = 1024 + 256 + 2048 + 64;   // alignment slack + barriers + per-CTA channel sums + ticket
  int stages = (int)((g_igemm_smem - overhead) / stage_bytes);
  if (stages < 2) stages = 2;
  if (stages > 8) stages = 8;
  const int kb_total = a->n_taps * a->cblocks;
  if (stages > kb_total) stages = kb_total < 2 ? 2 : kb_total;
  P.stages = stages;
  P.stage_tx = P.tile_w * P.tile_rows * 128 + P.b_stage_bytes;
  P.tmem_cols = P.n_tile <= 32 ? 32 : P.n_tile <= 64 ? 64 : P.n_tile <= 128 ? 128 : 256;
  P.out = reinterpret_cast<__nv_bfloat16*>(a->out);
  P.res = reinterpret_cast<const __nv_bfloat16*>(a->res);
  P.out_sn = a->out_sn; P.out_sh = a->out_sh; P.out_sw = a->out_sw;
  P.out_c = a->out_c;
  P.stat_sum = a->stat_sum;
  P.stat_sq = a->stat_sq;
  P.stat_replicas = a->stat_replicas > 1 ? a->stat_replicas : 1;
  P.stat_stride = a->stat_stride;
  P.fin_counter = nullptr;
  if (a->bn) {
    const mp_bn_branch* b = a->bn;
    MP_CHECK_ARG(a->stat_sum && P.stat_replicas == 1, "mp_conv_igemm: BatchNorm finalize needs un-replicated statistics");
    MP_CHECK_ARG(a->bn_counter && b->gamma && b->beta && b->save_mean && b->save_invstd && b->scale && b->shift &&
                     a->bn_count > 0 && a->bn_channels > 0 && a->bn_channels <= a->out_c,
                 "mp_conv_igemm: incomplete BatchNorm finalize arguments");
    P.fin_gamma = b->gamma; P.fin_beta = b->beta; P.fin_bias = b->conv_bias;
    P.fin_rmean = b->running_mean; P.fin_rvar = b->running_var;
    P.fin_smean = b->save_mean; P.fin_sinvstd = b->save_invstd; P.fin_scale = b->scale; P.fin_shift = b->shift;
    P.fin_counter = a->bn_counter;
    P.fin_C = a->bn_channels; P.fin_Cp = a->out_c;
    P.fin_count = a->bn_count; P.fin_momentum = a->bn_momentum; P.fin_eps = a->bn_eps;
    P.fin_total = a->bn_total_ctas > 0 ? a->bn_total_ctas
                                       : a->n_img * P.tiles_h * P.tiles_w * (int)(a->w_rows / P.n_tile);
  }

  {   // cluster size: M tiles must split evenly, weight-tile slices must stay 1024-byte aligned
    const long long m_tiles = (long long)a->n_img * P.tiles_h * P.tiles_w;
    int c = (int)g_igemm_cluster;
    while (c > 1 && (m_tiles % c != 0 || (P.n_tile / c) % 8 != 0 || P.n_tile % c != 0)) c /= 2;
    P.cluster = c < 1 ? 1 : c;
  }
  CUtensorMap tmA0, tmA1, tmB;
  const uint32_t boxA[5] = {64, (uint32_t)P.tile_w, 1, (uint32_t)P.tile_rows, 1};
  int rc = view_to_tmap(&tmA0, a->src[0], boxA, "mp_conv_igemm src[0]");
  if (rc != MP_OK) return rc;
  if (use_src1) {
    rc = view_to_tmap(&tmA1, a->src[1], boxA, "mp_conv_igemm src[1]");
    if (rc != MP_OK) return rc;
  } else {
    tmA1 = tmA0;
  }
  {
    MP_CHECK_ARG(mp_aligned16(a->wmat), "mp_conv_igemm: wmat not 16-byte aligned");
    const uint64_t dims[2] = {(uint64_t)a->w_k, (uint64_t)a->w_rows};
    const uint64_t strides[2] = {2, (uint64_t)a->w_k * 2};
    const uint32_t box[2] = {64, (uint32_t)(P.n_tile / P.cluster)};
    rc = tc::encode_tmap(&tmB, a->wmat, 2, dims, strides, box);
    if (rc != MP_OK) return rc;
  }

  const size_t smem = (size_t)stages * stage_bytes + overhead;
  if (!g_attr_set) {
    MP_CUDA(cudaFuncSetAttribute(igemm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    g_attr_set = true;
  }
  dim3 grid((unsigned)(a->n_img * P.tiles_h * P.tiles_w), (unsigned)(a->w_rows / P.n_tile));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = dim3(NTHREADS);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = (cudaStream_t)stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = P.cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  MP_CUDA(cudaLaunchKernelEx(&cfg, igemm_kernel, tmA0, tmA1, tmB, P));
  MP_CHECK_LAUNCH("mp_conv_igemm");
  return MP_OK;
}
