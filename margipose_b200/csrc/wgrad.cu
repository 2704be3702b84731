// Convolution weight gradient on tcgen05 (sm_100a):
//
//   dW[m, slot_t, n] += sum_pixels  A[pixel, m] * B_t[pixel + shift_t, n]
//
// (nn.Conv2d: A = dY, B = x;  nn.ConvTranspose2d: A = x, B = dY.)  The reduction runs over
// pixels, which is NOT the contiguous dimension of an NHWC tensor, so both operands are fed to
// the tensor core as MN-major tiles: a TMA box {64 ch, kp_w, 1, kp_rows, 1} lands
// [pixels][64 channels] lines of 128 bytes in shared memory and the UMMA descriptors walk it with
// K = pixel (8-pixel groups 1024 B apart) and M/N = channel (64-channel groups one box apart).
// Each CTA owns one 128-row slice of m, up to `T` taps (one TMEM accumulator of n_cols columns
// per tap, A tile shared by the taps) and a strided share of the pixel chunks (split-K); the
// epilogue adds its partial sums into the fp32 gradient with vector red.global.add.
//
// Replaces the cuDNN wgrad calls autograd issues for the nn.Conv2d / nn.ConvTranspose2d sites
// of /root/reference/src/margipose/models/margipose_model.py:33,67-68,73-74,79-82 (SURVEY.md
// section 2b, K3).  Roofline: tensor pipe; algorithmic flops = 2 * pixels * m * n * taps.
#include "tc.cuh"
#include "../../include/margipose_b200.h"

int mp_pick_tile(int out_h, int out_w, int max_pix, int* tile_w, int* tile_rows);

namespace {

constexpr int NTHREADS = 192;

struct WgradParams {
  mp_tap taps[MP_MAX_TAPS];
  int n_taps, T;
  int kp_w, kp_rows, chunks_w, chunks_h, chunks_total;
  int b_boxes, box_bytes, stage_bytes, stages, stage_tx, tmem_cols, ksteps;
  int m_real, n_real, n_cols, n_slots, n_off;
  int vec_ok;
  int dbg;   // experiment switches (tunable wgrad_dbg): 1 = no epilogue atomics, 2 = MMA twice, 4 = no MMA
  float* dw;
};

__global__ void __launch_bounds__(NTHREADS)
wgrad_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
             const __grid_constant__ WgradParams P) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = P.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * P.stage_bytes);
  uint64_t* empty = full + stages;
  uint64_t* tmem_full = empty + stages;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int tap0 = blockIdx.x * P.T;
  const int ntap = min(P.T, P.n_taps - tap0);
  const int m0 = blockIdx.z * 128;
  // pixel chunks blockIdx.y, blockIdx.y + gridDim.y, ...
  const int my_chunks = (P.chunks_total - (int)blockIdx.y + (int)gridDim.y - 1) / (int)gridDim.y;

  if (threadIdx.x == 0) {
    for (int s = 0; s < stages; ++s) {
      tc::mbar_init(&full[s], 1);
      tc::mbar_init(&empty[s], 1);
    }
    tc::mbar_init(tmem_full, 1);
    tc::mbar_fence_init();
    tc::prefetch_tmap(&tmA);
    tc::prefetch_tmap(&tmB);
  }
  if (warp == 1) tc::tmem_alloc(tmem_holder, P.tmem_cols);
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_holder;

  if (warp == 0) {
    if (lane == 0) {   // ---------------------------------------------------------- TMA producer
      int s = 0;
      uint32_t ph = 0;
      const int per_img = P.chunks_w * P.chunks_h;
      for (int i = 0; i < my_chunks; ++i) {
        const int chunk = blockIdx.y + i * gridDim.y;
        const int img = chunk / per_img;
        const int rem = chunk - img * per_img;
        const int ch = rem / P.chunks_w, cw = rem - ch * P.chunks_w;
        const int h0 = ch * P.kp_rows, w0 = cw * P.kp_w;
        uint8_t* st = smem + (size_t)s * P.stage_bytes;
        tc::mbar_wait(&empty[s], ph ^ 1);
        tc::mbar_arrive_expect_tx(&full[s], (uint32_t)((2 + ntap * P.b_boxes) * P.box_bytes));
        for (int j = 0; j < 2; ++j)
          tc::tma_load_5d(&tmA, &full[s], st + (size_t)j * P.box_bytes, m0 + j * 64, w0, 0, h0, img);
        for (int t = 0; t < ntap; ++t) {
          const mp_tap tap = P.taps[tap0 + t];
          for (int j = 0; j < P.b_boxes; ++j)
            tc::tma_load_5d(&tmB, &full[s], st + (size_t)(2 + t * P.b_boxes + j) * P.box_bytes,
                            tap.c0 + P.n_off + j * 64, w0 + tap.dw, tap.p, h0 + tap.dh, img);
        }
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {   // ------------------------------------------------------------ MMA issuer
      const uint32_t idesc = tc::idesc_bf16(128, P.n_cols, true, true);
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < my_chunks; ++i) {
        tc::mbar_wait(&full[s], ph);
        tc::tc_fence_after();
        const uint32_t st = tc::smem_u32(smem + (size_t)s * P.stage_bytes);
        for (int t = 0; t < ntap; ++t) {
          const uint32_t bbase = st + (uint32_t)((2 + t * P.b_boxes) * P.box_bytes);
          for (int rep = 0; rep < ((P.dbg & 2) ? 2 : 1); ++rep)
          for (int ks = 0; ks < P.ksteps; ++ks)   // 16 pixels = 16 lines of 128 B per step
            if (!(P.dbg & 4) || (i | ks) == 0)
            tc::mma_bf16(tmem + (uint32_t)(t * P.n_cols),
                         tc::desc_mnmajor_sw128(st + ks * 2048, P.box_bytes),
                         tc::desc_mnmajor_sw128(bbase + ks * 2048, P.box_bytes), idesc, (i | ks | rep) != 0);
        }
        tc::mma_commit(&empty[s]);
        if (++s == stages) { s = 0; ph ^= 1; }
      }
      tc::mma_commit(tmem_full);
    }
  } else {   // ------------------------------------------------------------------------ epilogue
    tc::mbar_wait(tmem_full, 0);
    tc::tc_fence_after();
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    const bool valid = row < P.m_real;
    for (int t = 0; t < ntap; ++t) {
      const int slot = P.taps[tap0 + t].koff;
      float* dst = P.dw + ((size_t)row * P.n_slots + slot) * P.n_real + P.n_off;
      const int n_left = P.n_real - P.n_off;
      for (int c = 0; c < P.n_cols / 32; ++c) {
        if (c * 32 >= n_left) break;   // warp-uniform
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * P.n_cols + c * 32), v);
        if (!valid || (P.dbg & 1)) continue;
        if (P.vec_ok) {
#pragma unroll
          for (int j = 0; j < 8; ++j)
            if (c * 32 + j * 4 < n_left)
              tc::red_add_v4(dst + c * 32 + j * 4, v[j * 4], v[j * 4 + 1], v[j * 4 + 2], v[j * 4 + 3]);
        } else {
#pragma unroll
          for (int j = 0; j < 32; ++j)
            if (c * 32 + j < n_left) atomicAdd(dst + c * 32 + j, v[j]);
        }
      }
    }
  }
  tc::tc_fence_before();
  __syncthreads();
  if (warp == 1) tc::tmem_dealloc(tmem, P.tmem_cols);
}

long long g_wgrad_ctas = 148;
long long g_wgrad_taps = 1;
long long g_wgrad_dbg = 0;
long long g_wgrad_kp = 128;
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

void mp_set_wgrad_tunable(int which, long long v) {
  if (which == 0) g_wgrad_ctas = v;
  else if (which == 1) g_wgrad_taps = v;
  else if (which == 2) g_wgrad_dbg = v;
  else g_wgrad_kp = v;
}

extern "C" int mp_conv_wgrad(const mp_wgrad_args* a, void* stream) {
  MP_CHECK_ARG(a, "mp_conv_wgrad: null args");
  MP_CHECK_ARG(a->n_taps >= 1 && a->n_taps <= MP_MAX_TAPS, "mp_conv_wgrad: n_taps %d out of range", a->n_taps);
  MP_CHECK_ARG(a->a.ptr && a->b.ptr && a->dw, "mp_conv_wgrad: null tensor");
  MP_CHECK_ARG(a->n_cols >= 64 && a->n_cols % 64 == 0 && a->n_cols <= 256,
               "mp_conv_wgrad: n_cols %d must be 64, 128, 192 or 256", a->n_cols);
  MP_CHECK_ARG(a->m_real > 0 && a->n_real > 0 && a->n_off >= 0 && a->n_off % 64 == 0 && a->n_off < a->n_real,
               "mp_conv_wgrad: bad channel counts");
  MP_CHECK_ARG(a->n_img > 0 && a->grid_h > 0 && a->grid_w > 0, "mp_conv_wgrad: empty pixel grid");
  for (int i = 0; i < a->n_taps; ++i)
    MP_CHECK_ARG(a->taps[i].koff >= 0 && a->taps[i].koff < a->n_slots, "mp_conv_wgrad: tap %d: bad slot", i);

  WgradParams P;
  for (int i = 0; i < a->n_taps; ++i) P.taps[i] = a->taps[i];
  P.n_taps = a->n_taps;
  int T = (int)g_wgrad_taps;
  if (T > 512 / a->n_cols) T = 512 / a->n_cols;
  if (T > a->n_taps) T = a->n_taps;
  if (T < 1) T = 1;
  const int groups = (a->n_taps + T - 1) / T;
  T = (a->n_taps + groups - 1) / groups;
  P.T = T;
  int kp_max = (int)g_wgrad_kp;
  // a stage holds (2 + T * n_cols/64) boxes of kp pixels x 128 B; keep at least 2 stages in 200 KB
  while (kp_max > 32 && 2 * (2 + T * (a->n_cols / 64)) * kp_max * 128 > 200 * 1024) kp_max /= 2;
  mp_pick_tile(a->grid_h, a->grid_w, kp_max, &P.kp_w, &P.kp_rows);
  while (P.kp_rows > 1 && (P.kp_w * P.kp_rows) % 16 != 0) --P.kp_rows;
  const int kp = P.kp_w * P.kp_rows;
  MP_CHECK_ARG(kp % 16 == 0, "mp_conv_wgrad: cannot form a 16-pixel-aligned chunk from a %dx%d grid",
               a->grid_h, a->grid_w);
  P.chunks_w = (a->grid_w + P.kp_w - 1) / P.kp_w;
  P.chunks_h = (a->grid_h + P.kp_rows - 1) / P.kp_rows;
  P.chunks_total = a->n_img * P.chunks_w * P.chunks_h;
  P.b_boxes = a->n_cols / 64;
  P.box_bytes = kp * 128;
  P.stage_bytes = (2 + T * P.b_boxes) * P.box_bytes;
  const int overhead = 1024 + 256;
  int stages = (200 * 1024 - overhead) / P.stage_bytes;
  if (stages < 2) stages = 2;
  if (stages > 8) stages = 8;
  P.stages = stages;
  P.ksteps = kp / 16;
  const int cols = T * a->n_cols;
  P.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
  P.m_real = a->m_real; P.n_real = a->n_real; P.n_cols = a->n_cols; P.n_slots = a->n_slots; P.n_off = a->n_off;
  P.vec_ok = (a->n_real % 4 == 0) && mp_aligned16(a->dw);
  P.dw = a->dw;
  P.dbg = (int)g_wgrad_dbg;

  const int m_tiles = (a->m_real + 127) / 128;
  int split = (int)(g_wgrad_ctas / (groups * m_tiles));
  if (split < 1) split = 1;
  if (split > P.chunks_total) split = P.chunks_total;

  CUtensorMap tmA, tmB;
  const uint32_t box[5] = {64, (uint32_t)P.kp_w, 1, (uint32_t)P.kp_rows, 1};
  int rc = view_to_tmap(&tmA, a->a, box, "mp_conv_wgrad a");
  if (rc != MP_OK) return rc;
  rc = view_to_tmap(&tmB, a->b, box, "mp_conv_wgrad b");
  if (rc != MP_OK) return rc;

  const size_t smem = (size_t)stages * P.stage_bytes + overhead;
  MP_CHECK_ARG(smem <= 227 * 1024, "mp_conv_wgrad: stage too large (%zu bytes of shared memory)", smem);
  if (!g_attr_set) {
    MP_CUDA(cudaFuncSetAttribute(wgrad_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    g_attr_set = true;
  }
  dim3 grid((unsigned)groups, (unsigned)split, (unsigned)m_tiles);
  wgrad_kernel<<<grid, NTHREADS, smem, (cudaStream_t)stream>>>(tmA, tmB, P);
  MP_CHECK_LAUNCH("mp_conv_wgrad");
  return MP_OK;
}
