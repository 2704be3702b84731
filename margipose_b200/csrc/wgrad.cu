// Convolution weight gradient on tcgen05 (sm_100a):
//
//   dW[m, slot_t, n] += sum_pixels  A[pixel, m] * B_t[pixel + shift_t, n]
//
// (nn.Conv2d: A = dY, B = x;  nn.ConvTranspose2d: A = x, B = dY.)  The reduction runs over
// pixels, which is NOT the contiguous dimension of an NHWC tensor, so both operands are fed to
// the tensor core as MN-major tiles: a TMA box {64 ch, kp_w, 1, kp_rows, 1} lands
// [pixels][64 channels] lines of 128 bytes in shared memory and the UMMA descriptors walk it with
// K = pixel (8-pixel groups 1024 B apart) and M/N = channel (64-channel groups one box apart).
// Each CTA owns one 128-row slice of m, one TAP GROUP, one slice of the n columns and a strided
// share of the pixel chunks (split-K).  By default a tap group is a single tap.  With the tunable
// wgrad_halo it is up to three taps that differ only by their row shift dh (a 3x3 filter column):
// their B tiles are row-shifted windows of ONE box of kp_rows + 2 rows, fetched once, each tap's
// descriptor starts dh * kp_w * 128 bytes further (a whole number of swizzle atoms) and the A tile
// is shared as well -- 2.4x less L2 -> shared-memory traffic for 3x3 filters, but three times the
// split-K atomics: measured 1.5 % slower in the training step, hence optional.  One TMEM accumulator
// per tap; the epilogue adds the partial sums into the fp32 gradient with vector red.global.add.
// A grouped launch (grid z = problem x m-tile) runs the three HeatmapColumns of a stage at once.
//
// Replaces the cuDNN wgrad calls autograd issues for the nn.Conv2d / nn.ConvTranspose2d sites
// of /root/reference/src/margipose/models/margipose_model.py:33,67-68,73-74,79-82 (SURVEY.md
// section 2b, K3).  Roofline: tensor pipe; algorithmic flops = 2 * pixels * m * n * taps.
#include <string.h>
#include "tc.cuh"
#include "../../include/margipose_b200.h"

int mp_pick_tile(int out_h, int out_w, int max_pix, int* tile_w, int* tile_rows);

namespace {

constexpr int NTHREADS = 192;

// taps (c0, dw, p) with row shifts dh0 .. dh0 + n - 1 (n > 1: one halo box of kp_rows + 2 rows)
struct TapGroup {
  int c0, dw, p, dh0, n;
  int slot[3];
};

struct WgradParams {
  TapGroup groups[MP_MAX_TAPS];
  int n_groups, n_slices;
  int slice_off[4], slice_cols[4];   // column slices of b handled by separate CTAs (TMEM: 3 taps x cols <= 512)
  int kp_w, kp_rows, chunks_w, chunks_h, chunks_total;
  int a_box_bytes, stage_bytes, stages, tmem_cols, ksteps, row_bytes;
  int m_real, n_real, n_slots, n_off;
  int vec_ok;
  int dbg;   // experiment switches (tunable wgrad_dbg): 1 = no epilogue atomics, 4 = no MMA
  int m_tiles;
  int n_pass;   // 1, or 3 in split (bf16x3) mode: every pixel chunk is accumulated as a_hi*b_hi + a_lo*b_hi + a_hi*b_lo
  float* dw[MP_MAX_GROUP];   // per problem of a grouped launch (blockIdx.z / m_tiles)
};

template <int NPASS>
struct alignas(64) WgradMaps {
  CUtensorMap a[MP_MAX_GROUP], b[MP_MAX_GROUP], bh[MP_MAX_GROUP];
};
template <>
struct alignas(64) WgradMaps<3> {
  CUtensorMap a[MP_MAX_GROUP], b[MP_MAX_GROUP], bh[MP_MAX_GROUP];
  CUtensorMap a2[MP_MAX_GROUP], b2[MP_MAX_GROUP], bh2[MP_MAX_GROUP];   // low halves of the operand pairs (split mode)
};

template <int NPASS>
__global__ void __launch_bounds__(NTHREADS)
wgrad_kernel(const __grid_constant__ WgradMaps<NPASS> TM, const __grid_constant__ WgradParams P) {
  const int prob = blockIdx.z / P.m_tiles;
  const CUtensorMap& tmA = TM.a[prob];
  const CUtensorMap& tmB = TM.b[prob];
  const CUtensorMap& tmBh = TM.bh[prob];
  const CUtensorMap* tmA2 = &tmA;
  const CUtensorMap* tmB2 = &tmB;
  const CUtensorMap* tmBh2 = &tmBh;
  if constexpr (NPASS == 3) {
    tmA2 = &TM.a2[prob];
    tmB2 = &TM.b2[prob];
    tmBh2 = &TM.bh2[prob];
  }
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  const int stages = P.stages;
  uint64_t* full = reinterpret_cast<uint64_t*>(smem + (size_t)stages * P.stage_bytes);
  uint64_t* empty = full + stages;
  uint64_t* tmem_full = empty + stages;
  uint32_t* tmem_holder = reinterpret_cast<uint32_t*>(tmem_full + 1);

  const int warp = tc::warp_idx_uniform(), lane = threadIdx.x & 31;
  const TapGroup& G = P.groups[blockIdx.x / P.n_slices];
  const int slice = blockIdx.x % P.n_slices;
  const int n_off = P.n_off + P.slice_off[slice], n_cols = P.slice_cols[slice];
  const int b_boxes = n_cols / 64;
  const bool halo = G.n > 1;
  const int b_box_bytes = halo ? (P.kp_rows + 2) * P.row_bytes : P.a_box_bytes;
  const int m0 = (blockIdx.z - prob * P.m_tiles) * 128;
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
    tc::prefetch_tmap(&tmBh);
  }
  if (warp == 1) tc::tmem_alloc(tmem_holder, P.tmem_cols);
  pdl_trigger();
  tc::tc_fence_before();
  __syncthreads();
  tc::tc_fence_after();
  const uint32_t tmem = *tmem_holder;
  pdl_wait();   // everything above overlapped the previous kernel's tail; its results are visible from here on

  if (warp == 0) {
    {   // --------------------------------------------- TMA producer: warp-uniform loop, one elected lane issues
      int s = 0;
      uint32_t ph = 0;
      const int per_img = P.chunks_w * P.chunks_h;
      const uint32_t tx = (uint32_t)(2 * P.a_box_bytes + (halo ? 1 : G.n) * b_boxes * b_box_bytes);
      for (int it = 0; it < my_chunks * NPASS; ++it) {
        const int i = it / NPASS, pass = it - i * NPASS;   // pass 1 reads a_lo, pass 2 reads b_lo
        const CUtensorMap* mA = pass == 1 ? tmA2 : &tmA;
        const CUtensorMap* mB = pass == 2 ? (halo ? tmBh2 : tmB2) : (halo ? &tmBh : &tmB);
        const int chunk = blockIdx.y + i * gridDim.y;
        const int img = chunk / per_img;
        const int rem = chunk - img * per_img;
        const int ch = rem / P.chunks_w, cw = rem - ch * P.chunks_w;
        const int h0 = ch * P.kp_rows, w0 = cw * P.kp_w;
        uint8_t* st = smem + (size_t)s * P.stage_bytes;
        tc::mbar_wait(&empty[s], ph ^ 1);
        if (tc::elect_one()) {
          tc::mbar_arrive_expect_tx(&full[s], tx);
          for (int j = 0; j < 2; ++j)
            tc::tma_load_5d(mA, &full[s], st + (size_t)j * P.a_box_bytes, m0 + j * 64, w0, 0, h0, img);
          uint8_t* sb = st + 2 * (size_t)P.a_box_bytes;
          for (int j = 0; j < b_boxes; ++j)
            tc::tma_load_5d(mB, &full[s], sb + (size_t)j * b_box_bytes, G.c0 + n_off + j * 64,
                            w0 + G.dw, G.p, h0 + G.dh0, img);
        }
        if (++s == stages) { s = 0; ph ^= 1; }
      }
    }
  } else if (warp == 1) {
    {   // ----------------------------------------------- MMA issuer: warp-uniform loop, one elected lane issues
      const uint32_t idesc = tc::idesc_bf16(128, n_cols, true, true);
      int s = 0;
      uint32_t ph = 0;
      for (int i = 0; i < my_chunks * NPASS; ++i) {
        tc::mbar_wait(&full[s], ph);
        tc::tc_fence_after();
        const uint32_t st = tc::smem_u32(smem + (size_t)s * P.stage_bytes);
        const uint32_t sb = st + 2u * (uint32_t)P.a_box_bytes;
        if (tc::elect_one()) {
          // 16 pixels = 16 lines of 128 B per K step (2048 B -> +128 in the descriptor's address field); tap t reads the
          // row-shifted window t of the halo box.  (Taking turns between the taps' accumulators per K step measured
          // slower here than finishing one tap's K steps first.)
          const uint64_t ad = tc::desc_mnmajor_sw128(st, P.a_box_bytes);
          const uint64_t bd0 = tc::desc_mnmajor_sw128(sb, b_box_bytes);
          const uint32_t tap_step = (uint32_t)P.row_bytes >> 4;
          for (int t = 0; t < G.n; ++t)
            for (int ks = 0; ks < P.ksteps; ++ks)
              if (!(P.dbg & 4) || (i | ks) == 0)
                tc::mma_bf16(tmem + (uint32_t)(t * n_cols), ad + 128 * ks, bd0 + (uint64_t)(t * tap_step) + 128 * ks, idesc,
                             (i | ks) != 0);
          tc::mma_commit(&empty[s]);
        }
        if (++s == stages) { s = 0; ph ^= 1; }
      }
      if (tc::elect_one()) tc::mma_commit(tmem_full);
    }
  } else {   // ------------------------------------------------------------------------ epilogue
    tc::mbar_wait(tmem_full, 0);
    tc::tc_fence_after();
    const int q = warp & 3;
    const int row = m0 + q * 32 + lane;
    const bool valid = row < P.m_real;
    for (int t = 0; t < G.n; ++t) {
      float* dst = P.dw[prob] + ((size_t)row * P.n_slots + G.slot[t]) * P.n_real + n_off;
      const int n_left = P.n_real - n_off;
      for (int c = 0; c < n_cols / 32; ++c) {
        if (c * 32 >= n_left) break;   // warp-uniform
        float v[32];
        tc::tmem_ld32(tmem + ((uint32_t)(q * 32) << 16) + (uint32_t)(t * n_cols + c * 32), v);
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
long long g_wgrad_halo = 0;   // measured in the grouped step: one tap per CTA (more split-K CTAs, 3 stages) is 1.5 % faster
long long g_wgrad_dbg = 0;
long long g_wgrad_kp = 128;
long long g_wgrad_slice = 256;
long long g_wgrad_smem = 200 * 1024;   // shared-memory budget for the pipeline stages
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

// Groups of up to `max_n` taps sharing (c0, dw, p) with consecutive row shifts; the rest stay single.
int group_taps(const mp_wgrad_args* a, int max_n, TapGroup* out) {
  bool used[MP_MAX_TAPS] = {};
  int n_groups = 0;
  for (int i = 0; i < a->n_taps; ++i) {
    if (used[i]) continue;
    const mp_tap& t = a->taps[i];
    used[i] = true;
    int dh[3] = {t.dh, 0, 0}, slot[3] = {t.koff, 0, 0}, n = 1;
    int lo = t.dh, hi = t.dh;
    bool grew = true;
    while (grew && n < max_n) {
      grew = false;
      for (int j = 0; j < a->n_taps && n < max_n; ++j) {
        const mp_tap& u = a->taps[j];
        if (used[j] || u.c0 != t.c0 || u.dw != t.dw || u.p != t.p) continue;
        if (u.dh != hi + 1 && u.dh != lo - 1) continue;
        if (u.dh > hi) hi = u.dh; else lo = u.dh;
        dh[n] = u.dh; slot[n] = u.koff; ++n;
        used[j] = true;
        grew = true;
      }
    }
    TapGroup G;
    G.c0 = t.c0; G.dw = t.dw; G.p = t.p; G.dh0 = lo; G.n = n;
    G.slot[0] = G.slot[1] = G.slot[2] = 0;
    for (int k = 0; k < n; ++k) G.slot[dh[k] - lo] = slot[k];
    out[n_groups++] = G;
  }
  return n_groups;
}

}  // namespace

void mp_set_wgrad_tunable(int which, long long v) {
  if (which == 0) g_wgrad_ctas = v;
  else if (which == 1) g_wgrad_halo = v;
  else if (which == 2) g_wgrad_dbg = v;
  else if (which == 3) g_wgrad_kp = v;
  else if (which == 4) g_wgrad_slice = v;
  else g_wgrad_smem = v;
}

static int check_one(const mp_wgrad_args* a) {
  MP_CHECK_ARG(a->n_taps >= 1 && a->n_taps <= MP_MAX_TAPS, "mp_conv_wgrad: n_taps %d out of range", a->n_taps);
  MP_CHECK_ARG(a->a.ptr && a->b.ptr && a->dw, "mp_conv_wgrad: null tensor");
  MP_CHECK_ARG((a->a_lo.ptr == nullptr) == (a->b_lo.ptr == nullptr), "mp_conv_wgrad: a_lo and b_lo go together");
  MP_CHECK_ARG(a->n_cols >= 64 && a->n_cols % 64 == 0 && a->n_cols <= 256,
               "mp_conv_wgrad: n_cols %d must be 64, 128, 192 or 256", a->n_cols);
  MP_CHECK_ARG(a->m_real > 0 && a->n_real > 0 && a->n_off >= 0 && a->n_off % 64 == 0 && a->n_off < a->n_real,
               "mp_conv_wgrad: bad channel counts");
  MP_CHECK_ARG(a->n_img > 0 && a->grid_h > 0 && a->grid_w > 0, "mp_conv_wgrad: empty pixel grid");
  for (int i = 0; i < a->n_taps; ++i)
    MP_CHECK_ARG(a->taps[i].koff >= 0 && a->taps[i].koff < a->n_slots, "mp_conv_wgrad: tap %d: bad slot", i);
  return MP_OK;
}

static bool same_view(const mp_view5& x, const mp_view5& y) {
  for (int i = 0; i < 5; ++i)
    if (x.dim[i] != y.dim[i] || x.stride[i] != y.stride[i]) return false;
  return true;
}

static bool same_geometry(const mp_wgrad_args* x, const mp_wgrad_args* y) {
  return x->n_taps == y->n_taps && x->m_real == y->m_real && x->n_real == y->n_real && x->n_cols == y->n_cols &&
         x->n_slots == y->n_slots && x->n_img == y->n_img && x->grid_h == y->grid_h && x->grid_w == y->grid_w &&
         x->n_off == y->n_off && same_view(x->a, y->a) && same_view(x->b, y->b) &&
         (x->a_lo.ptr == nullptr) == (y->a_lo.ptr == nullptr) &&
         (((uintptr_t)x->dw ^ (uintptr_t)y->dw) & 15) == 0 && memcmp(x->taps, y->taps, sizeof(mp_tap) * x->n_taps) == 0;
}

extern "C" int mp_conv_wgrad(const mp_wgrad_args* a, void* stream) { return mp_conv_wgrad_grouped(a, 1, stream); }

extern "C" int mp_conv_wgrad_grouped(const mp_wgrad_args* args, int n_problems, void* stream) {
  MP_CHECK_ARG(args, "mp_conv_wgrad: null args");
  MP_CHECK_ARG(n_problems >= 1 && n_problems <= MP_MAX_GROUP, "mp_conv_wgrad_grouped: %d problems (1..%d)", n_problems,
               MP_MAX_GROUP);
  const mp_wgrad_args* a = &args[0];
  for (int i = 0; i < n_problems; ++i) {
    int rc = check_one(&args[i]);
    if (rc != MP_OK) return rc;
    MP_CHECK_ARG(i == 0 || same_geometry(a, &args[i]), "mp_conv_wgrad_grouped: problem %d differs in geometry", i);
  }

  WgradParams P;
  int kp_max = (int)g_wgrad_kp;
  mp_pick_tile(a->grid_h, a->grid_w, kp_max, &P.kp_w, &P.kp_rows);
  while (P.kp_rows > 1 && (P.kp_w * P.kp_rows) % 16 != 0) --P.kp_rows;
  int kp = P.kp_w * P.kp_rows;
  MP_CHECK_ARG(kp % 16 == 0, "mp_conv_wgrad: cannot form a 16-pixel-aligned chunk from a %dx%d grid",
               a->grid_h, a->grid_w);
  P.row_bytes = P.kp_w * 128;

  // tap groups (halo boxes need whole swizzle atoms per pixel row) and the column slices that keep
  // 3 accumulators within the 512 TMEM columns: 192 -> 128 + 64, 256 -> 128 + 128
  const bool allow_halo = g_wgrad_halo != 0 && P.kp_w % 8 == 0 && a->n_taps > 1;
  P.n_groups = group_taps(a, allow_halo ? 3 : 1, P.groups);
  int max_n = 1;
  for (int g = 0; g < P.n_groups; ++g) max_n = P.groups[g].n > max_n ? P.groups[g].n : max_n;
  P.n_slices = 1;
  for (int i = 0; i < 4; ++i) { P.slice_off[i] = 0; P.slice_cols[i] = 0; }
  P.slice_cols[0] = a->n_cols;
  {
    int width = a->n_cols;   // widest slice: 3 accumulators must fit TMEM; the tunable can force narrower ones
    if (max_n * width > 512) width = 128;
    if (max_n > 1 && g_wgrad_slice >= 64 && g_wgrad_slice < width) width = (int)g_wgrad_slice / 64 * 64;
    if (width < a->n_cols) {
      P.n_slices = 0;
      for (int off = 0; off < a->n_cols; off += width) {
        P.slice_off[P.n_slices] = off;
        P.slice_cols[P.n_slices] = a->n_cols - off < width ? a->n_cols - off : width;
        ++P.n_slices;
      }
    }
  }
  const int max_cols = P.slice_cols[0];
  P.a_box_bytes = kp * 128;
  const int b_box_plain = kp * 128, b_box_halo = (P.kp_rows + 2) * P.row_bytes;
  P.stage_bytes = 2 * P.a_box_bytes + (max_cols / 64) * (max_n > 1 ? b_box_halo : b_box_plain);
  const int overhead = 1024 + 256;
  int stages = (int)((g_wgrad_smem - overhead) / P.stage_bytes);
  MP_CHECK_ARG(stages >= 1, "mp_conv_wgrad: stage of %d bytes does not fit", P.stage_bytes);
  if (stages < 2) stages = 2;
  if (stages > 8) stages = 8;
  P.chunks_w = (a->grid_w + P.kp_w - 1) / P.kp_w;
  P.chunks_h = (a->grid_h + P.kp_rows - 1) / P.kp_rows;
  P.chunks_total = a->n_img * P.chunks_w * P.chunks_h;
  P.stages = stages;
  P.ksteps = kp / 16;
  const int cols = max_n * max_cols;
  P.tmem_cols = cols <= 32 ? 32 : cols <= 64 ? 64 : cols <= 128 ? 128 : cols <= 256 ? 256 : 512;
  P.m_real = a->m_real; P.n_real = a->n_real; P.n_slots = a->n_slots; P.n_off = a->n_off;
  P.vec_ok = (a->n_real % 4 == 0) && mp_aligned16(a->dw);
  P.dbg = (int)g_wgrad_dbg;

  const int m_tiles = (a->m_real + 127) / 128;
  P.m_tiles = m_tiles;
  P.n_pass = a->a_lo.ptr ? 3 : 1;
  const int gx = P.n_groups * P.n_slices;
  int split = (int)(g_wgrad_ctas / (gx * m_tiles * n_problems));
  if (split < 1) split = 1;
  if (split > P.chunks_total) split = P.chunks_total;
  // deterministic mode: no split-K, every gradient element receives exactly one add per launch (launches that
  // accumulate into the same tensor are ordered by their stream)
  if (mp_deterministic()) split = 1;

  WgradMaps<3> TM;   // (the single-pass kernel receives the first three arrays only)
  const uint32_t box[5] = {64, (uint32_t)P.kp_w, 1, (uint32_t)P.kp_rows, 1};
  const uint32_t boxh[5] = {64, (uint32_t)P.kp_w, 1, (uint32_t)(P.kp_rows + 2), 1};
  for (int i = 0; i < MP_MAX_GROUP; ++i) {
    const mp_wgrad_args* x = &args[i < n_problems ? i : 0];
    P.dw[i] = x->dw;
    int rc = view_to_tmap(&TM.a[i], x->a, box, "mp_conv_wgrad a");
    if (rc != MP_OK) return rc;
    rc = view_to_tmap(&TM.b[i], x->b, box, "mp_conv_wgrad b");
    if (rc != MP_OK) return rc;
    if (max_n > 1) {
      rc = view_to_tmap(&TM.bh[i], x->b, boxh, "mp_conv_wgrad b (halo box)");
      if (rc != MP_OK) return rc;
    } else {
      TM.bh[i] = TM.b[i];
    }
    TM.a2[i] = TM.a[i]; TM.b2[i] = TM.b[i]; TM.bh2[i] = TM.bh[i];
    if (P.n_pass > 1) {
      rc = view_to_tmap(&TM.a2[i], x->a_lo, box, "mp_conv_wgrad a_lo");
      if (rc != MP_OK) return rc;
      rc = view_to_tmap(&TM.b2[i], x->b_lo, box, "mp_conv_wgrad b_lo");
      if (rc != MP_OK) return rc;
      TM.bh2[i] = TM.b2[i];
      if (max_n > 1) {
        rc = view_to_tmap(&TM.bh2[i], x->b_lo, boxh, "mp_conv_wgrad b_lo (halo box)");
        if (rc != MP_OK) return rc;
      }
    }
  }

  const size_t smem = (size_t)stages * P.stage_bytes + overhead;
  MP_CHECK_ARG(smem <= 227 * 1024, "mp_conv_wgrad: stage too large (%zu bytes of shared memory)", smem);
  if (!g_attr_set) {
    MP_CUDA(cudaFuncSetAttribute(wgrad_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    MP_CUDA(cudaFuncSetAttribute(wgrad_kernel<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    g_attr_set = true;
  }
  dim3 grid((unsigned)gx, (unsigned)split, (unsigned)(m_tiles * n_problems));
  if (P.n_pass > 1) {
    MP_CUDA(mp_launch(wgrad_kernel<3>, grid, dim3(NTHREADS), smem, (cudaStream_t)stream, TM, P));
  } else {
    WgradMaps<1> T1;
    memcpy(&T1, &TM, sizeof(T1));
    MP_CUDA(mp_launch(wgrad_kernel<1>, grid, dim3(NTHREADS), smem, (cudaStream_t)stream, T1, P));
  }
  MP_CHECK_LAUNCH("mp_conv_wgrad");
  return MP_OK;
}
