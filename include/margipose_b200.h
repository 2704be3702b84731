/* margipose_b200 -- C ABI of the B200-native MargiPose hot path.
 *
 * The reference (anibali/margipose) is pure Python/PyTorch and has no FFI; its extension
 * point is the Python model registry (src/margipose/models/__init__.py:10-27).  This header
 * is the drop-in boundary underneath the Python mirror of that registry
 * (margipose_b200/models): every entry point replaces a group of PyTorch op sites on the
 * reference's hot path (cited per function; paths relative to /root/reference/src/margipose/).
 *
 * Conventions
 *  - plain pointers and sizes only; every pointer is a DEVICE pointer unless stated;
 *  - the caller owns and allocates every buffer, including workspaces; no hidden allocation;
 *  - work is enqueued on `stream` (a cudaStream_t passed as void*); no implicit sync;
 *  - returns 0 on success, a negative MP_ERR_* code otherwise; never throws;
 *    mp_last_error() returns a thread-local description of the last failure;
 *  - tensors: heatmaps / logits are fp32 NCHW-contiguous (B, J, H, W) as in the reference;
 *    network activations are bf16 NHWC (see DESIGN.md "Data layout");
 *  - `_grouped` entry points run 1..MP_MAX_GROUP problems of identical geometry (the xy / zy / xz
 *    HeatmapColumns of a stage) in one launch; the plain entry point is the 1-problem case;
 *  - the kernels of the network chain are launched with programmatic stream serialization (tunable "pdl"):
 *    on one stream a kernel's prologue may overlap its predecessor's tail, results are identical.
 */
#ifndef MARGIPOSE_B200_H
#define MARGIPOSE_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#if defined(__GNUC__)
#define MP_API __attribute__((visibility("default")))
#else
#define MP_API
#endif

#define MP_OK 0
#define MP_ERR_ARG (-1)
#define MP_ERR_CUDA (-2)
#define MP_ERR_UNSUPPORTED (-3)

/* ABI version of this header; bumped on any signature change. */
MP_API int mp_abi_version(void);
/* Thread-local message for the last non-zero return code. Host pointer, never NULL. */
MP_API const char* mp_last_error(void);

/* ---------------------------------------------------------------------------------------
 * Fused tail.
 *
 * mp_tail_fwd: for each (b, j) and each non-NULL plane k in {xy, zy, xz}:
 *   p_k = softmax(in_k) over H*W if from_logits else in_k          dsntnn.py:124-130
 *   (a_k, b_k) = (E[col coord], E[row coord]) of p_k               dsntnn.py:39-62,84-96
 *   js_k = JS(p_k || normalised Gaussian(mu_k, sigma px))          dsntnn.py:154-232
 *   coords = (a_xy, b_xy, (a_zy + b_xz)/2)                         models/margipose_model.py:254-261
 *   loss  = js_xy + js_zy + js_xz + |coords - target|              models/margipose_model.py:236-252
 *          (2D samples, valid_depth[b]==0: js_xy + |coords.xy - target.xy|, :223-234)
 * Outputs are all optional (NULL skips): prob[k] (B,J,H,W); ab[k] (B,J,2); js[k] (B,J);
 * coords (B,J,3); loss (B,J) (added to the existing value when accumulate != 0, which is how
 * the per-stage sum of the reference's loss loop is formed).
 * Targets: `target` (B,J,3) xyz drives the fused loss (per-plane means are derived as
 * xy->(x,y), zy->(z,y), xz->(x,z)); alternatively mu[k] (B,J,2) gives an explicit (col,row)
 * mean for plane k (generic js_reg_losses).  valid_depth: (B) int32 or NULL (= all 3D).
 * pixelwise: 1 = 'jsd', 0 = None (models/margipose_model.py:215-221).
 */
MP_API int mp_tail_fwd(const float* const in[3], int from_logits, float* const prob[3],
                float* const ab[3], float* const js[3], const float* const mu[3],
                const float* target, const int* valid_depth, float* coords, float* loss,
                int accumulate, int pixelwise, double sigma, int B, int J, int H, int W,
                void* stream);

/* mp_tail_bwd: gradient of the tail w.r.t. its input planes (replaces the ~440-node autograd
 * graph of one stage's tail).  For each plane k with out[k] != NULL:
 *   D = gup_k + w_js * dJS/dp + c_col * col_coord + c_row * row_coord
 *   out_k = project ? p_k * (D - sum(p_k * D)) : D        (project = softmax backward)
 * Fused mode (target != NULL): coefficients come from saved `coords` (B,J,3), `target` and
 * w (B,J) = dL/dloss[b,j].  Generic mode: coef[k] (B,J,3) = (w_js, c_col, c_row), with mu[k]
 * (B,J,2) required when w_js is used.  gup[k] (upstream gradient on p_k) may be NULL. */
MP_API int mp_tail_bwd(const float* const prob[3], const float* const gup[3], float* const out[3],
                const float* target, const float* coords, const float* w,
                const int* valid_depth, const float* const mu[3], const float* const coef[3],
                int project, int pixelwise, double sigma, int B, int J, int H, int W,
                void* stream);

/* average_loss (dsntnn.py:99-121): out2[0] = sum(l*mask)/max(sum(mask),1), out2[1] = denominator.
 * mask may be NULL (= ones). */
MP_API int mp_masked_mean_fwd(const float* losses, const float* mask, int n, float* out2, void* stream);
MP_API int mp_masked_mean_bwd(const float* grad_out, const float* mask, const float* mean_den, int n,
                       float* grad_losses, void* stream);

/* euclidean_losses (dsntnn.py:133-151) over n points of d dims, and its gradient. */
MP_API int mp_euclid_fwd(const float* actual, const float* target, int n, int d, float* out, void* stream);
MP_API int mp_euclid_bwd(const float* grad_out, const float* actual, const float* target,
                  const float* dist, int n, int d, float* grad_actual, void* stream);

/* make_gauss (dsntnn.py:154-195), 2D: mu (BJ,2) = (col,row) means in normalised units,
 * out (BJ,H,W); sigma in pixels. */
MP_API int mp_make_gauss(const float* mu, float* out, int normalize, double sigma, int BJ, int H, int W,
                  void* stream);

/* ---------------------------------------------------------------------------------------
 * Convolutions as implicit GEMM on the 5th-gen tensor cores.
 *
 * Replaces every nn.Conv2d / nn.ConvTranspose2d op site of models/margipose_model.py:33,67-68,
 * 73-74,79-82,125,145 and of the truncated torchvision ResNet (:130-135), forward and backward.
 *
 * Activations are bf16 NHWC with the channel count padded to a multiple of 64.  A source tensor
 * is handed over as a 5-D view (innermost first: C, W, P, H, N) so that one TMA box
 * {64 channels, tile_w, 1, tile_rows, 1} fetches the operand tile of a filter tap already
 * shifted (zero padding = TMA out-of-bounds fill) and, for stride-2 taps, already decimated:
 *   stride 1:  dims (C, W, 1, H, N)
 *   stride 2:  dims (2C, W/2, 2, H/2, N) over the same memory; a tap picks the column parity
 *              through c0 (0 or C) and the row parity through p.
 */
#define MP_MAX_TAPS 30   /* a 3x3 conv in split (bf16x3) mode: 9 taps x (hi*hi, lo*hi, hi*lo) */
#define MP_MAX_GROUP 3   /* problems per grouped launch (the three HeatmapColumns of a stage) */

typedef struct mp_view5 {
  const void* ptr;       /* bf16 */
  int64_t dim[5];        /* elements, innermost first */
  int64_t stride[5];     /* elements; stride[0] == 1; (stride * 2 bytes) % 16 == 0 */
} mp_view5;

typedef struct mp_tap {
  int32_t src;           /* source view index (0 or 1) */
  int32_t c0;            /* coordinate offset in dim 0 */
  int32_t dw;            /* added to the tile's w origin (dim 1) */
  int32_t p;             /* coordinate in dim 2 */
  int32_t dh;            /* added to the tile's h origin (dim 3) */
  int32_t koff;          /* igemm: first K column of this tap in wmat; wgrad: tap slot in dw */
} mp_tap;

/* out[n,h,w,:] (+= res[n,h,w,:]) = sum_taps sum_c src[tap](n, h+dh, w+dw, c) * wmat[:, koff + c]
 * over an (n_img, out_h, out_w) output pixel grid.  wmat is bf16 [w_rows][w_k] row-major with
 * w_rows % 32 == 0 (zero rows beyond the real Cout).  The output pixel (n,h,w) lives at
 * out + n*out_sn + h*out_sh + w*out_sw (elements; lets a transposed conv scatter one output
 * parity class per launch); out_c channels are stored (multiple of 8).  When stat_sum/stat_sq are
 * given, the per-channel sum and sum of squares of the stored (bf16-rounded) values are
 * atomically added to them -- the BatchNorm batch statistics of nn.BatchNorm2d in train mode
 * (optionally spread over stat_replicas copies that the BatchNorm kernels add up). */
struct mp_bn_branch;

typedef struct mp_igemm_args {
  mp_view5 src[2];
  const void* wmat;
  int64_t w_rows, w_k;
  int32_t n_taps;
  mp_tap taps[MP_MAX_TAPS];
  int32_t cblocks;                 /* 64-channel blocks per tap */
  int32_t n_img, out_h, out_w;
  void* out;                       /* bf16 */
  const void* res;                 /* bf16, same addressing as out, or NULL */
  int64_t out_sn, out_sh, out_sw;
  int32_t out_c;
  float* stat_sum;
  float* stat_sq;
  int32_t stat_replicas;           /* >= 1: CTA b adds into replica (b % stat_replicas) ...          */
  int64_t stat_stride;             /* ... at stat_sum/stat_sq + replica * stat_stride (spreads atomics) */
  /* BatchNorm finalize fused into the conv: the LAST CTA to add its statistics (ticket counter; the
   * statistics of one BatchNorm input may come from bn_launches launches of identical geometry, e.g.
   * the 4 output-parity classes of a transposed conv) turns sum/sq into mean / 1/std / scale / shift,
   * saves them and updates the running buffers -- so the BatchNorm forward kernels start with one small
   * load instead of a reduction.  bn: HOST pointer to the branch (gamma, beta, running_*, save_*, scale,
   * shift, conv_bias are used) or NULL.  Requires stat_replicas == 1. */
  const struct mp_bn_branch* bn;
  uint32_t* bn_counter;            /* device counter, zero before the first contributing launch */
  int32_t bn_launches;             /* launches (of this geometry) that share bn_counter; 0 or 1 = this one only */
  int32_t bn_channels;             /* real channel count C */
  int64_t bn_count;                /* elements per channel (N*H*W of the BatchNorm input) */
  float bn_momentum, bn_eps;
  /* Epilogue affine (inference: eval-mode BatchNorm folded into the conv, bin/infer_single.py:58-66):
   *   out = relu2?( relu1?(acc * ep_scale[c] + ep_shift[c]) + res )
   * ep_scale / ep_shift: fp32, w_rows entries (zeros beyond the real channels), 16-byte aligned, or both NULL;
   * ep_relu: 0 = no ReLU, 1 = ReLU before the residual add (MargiPose ResidualBlock), 2 = after it (ResNet blocks). */
  const float* ep_scale;
  const float* ep_shift;
  int32_t ep_relu;
  /* Split ("bf16x3") precision mode, lo_delta > 0: every activation VALUE is a pair of bf16 tensors, hi = bf16(v) and
   * lo = bf16(v - hi) stored lo_delta ELEMENTS after hi (~16 significant bits; see DESIGN.md "Precision modes").
   * out, res and acc_in are such pairs.  A convolution accumulates x_hi * W_hi + x_lo * W_hi + x_hi * W_lo: in ONE launch
   * when the tripled tap list fits MP_MAX_TAPS (src[0] = x_hi, src[1] = x_lo, wmat = [rows][K_hi | K_lo]), otherwise in
   * three launches chained through acc_in; the epilogue computes
   *   out = relu2?( relu1?((acc + acc_in) * scale + shift) + res )  on fp32 values and stores the pair.  The BatchNorm statistics are those of the stored pair sums.  lo_delta == 0: plain bf16. */
  const void* acc_in;
  int64_t lo_delta;
} mp_igemm_args;

MP_API int mp_conv_igemm(const mp_igemm_args* args, void* stream);
/* n_problems (1..MP_MAX_GROUP) convolutions of identical geometry (taps, shapes, strides, flags) on different
 * tensors in ONE launch (grid z = problem): the xy / zy / xz HeatmapColumns of a MargiPose stage
 * (models/margipose_model.py:196-198) are three such problems at every layer.  args: contiguous array. */
MP_API int mp_conv_igemm_grouped(const mp_igemm_args* args, int n_problems, void* stream);

/* Weight gradient: dw[m][slot][n] += sum over the (n_img, grid_h, grid_w) pixel grid of
 *   a(pix, m) * b[tap](pix + shift, n)
 * a: natural view (C, W, 1, H, N) of the tensor indexed by the GEMM row m (dY for nn.Conv2d,
 * x for nn.ConvTranspose2d); b: view addressed through the taps like in mp_conv_igemm.
 * dw is fp32 [m_real][n_slots][n_real] (the channels-last memory of the torch parameter).
 * One launch covers the n_cols (64..256, multiple of 64) channels of b starting at n_off. */
typedef struct mp_wgrad_args {
  mp_view5 a;
  mp_view5 b;
  int32_t n_taps;
  mp_tap taps[MP_MAX_TAPS];
  int32_t m_real, n_real, n_cols, n_slots;
  int32_t n_img, grid_h, grid_w;
  int32_t n_off;
  float* dw;
  /* Split (bf16x3) mode: low halves of the two operand pairs, same view geometry as a / b (ptr NULL = plain bf16).  The
   * launch then accumulates a * b + a_lo * b + a * b_lo per pixel chunk into the same accumulator (one flush). */
  mp_view5 a_lo;
  mp_view5 b_lo;
} mp_wgrad_args;

MP_API int mp_conv_wgrad(const mp_wgrad_args* args, void* stream);
/* n_problems weight gradients of identical geometry in one launch (see mp_conv_igemm_grouped). */
MP_API int mp_conv_wgrad_grouped(const mp_wgrad_args* args, int n_problems, void* stream);

/* ---------------------------------------------------------------------------------------
 * BatchNorm + activation + residual, forward and backward (nn.BatchNorm2d / nn.ReLU / `+` in
 * ResidualBlock, models/margipose_model.py:31-40, and in the torchvision ResNet blocks).
 *
 * One conv output `y` (bf16 (M, Cp), M = N*H*W pixels) with its BatchNorm is a "branch".  The
 * forward computes, per pixel and channel,
 *     out = post( pre(bn_a(y_a)) + [ bn_b(y_b) | res | 0 ] ),   pre/post = ReLU or identity
 * which covers relu(bn(y)) (first half of a block), relu(bn(y2)) + bn(ys) (MargiPose
 * ResidualBlock output), relu(bn(y2) + x) and relu(bn(y2) + bn(yd)) (ResNet blocks).
 * In training mode the batch statistics come from the per-channel sums the conv epilogue
 * accumulated (sum, sq), running statistics are updated with `momentum` (unbiased variance),
 * and mean / 1/std are saved for the backward; in eval mode the running statistics are used.
 */
typedef struct mp_bn_branch {
  const void* y;            /* bf16 (M, Cp); NULL = branch absent */
  const float* sum;         /* (Cp) batch sum of y        (training forward) */
  const float* sq;          /* (Cp) batch sum of y*y      (training forward) */
  const float* gamma;       /* (C) */
  const float* beta;        /* (C) */
  float* running_mean;      /* (C) */
  float* running_var;       /* (C) */
  float* save_mean;         /* (Cp) written by forward, read by backward */
  float* save_invstd;       /* (Cp) */
  const float* conv_bias;   /* (C) bias of the producing conv or NULL (only shifts the mean) */
  float* scale;             /* (Cp) gamma/std, written by the producing conv's finalize (see mp_igemm_args.bn) */
  float* shift;             /* (Cp) beta - mean*scale; when scale/shift are NULL the BatchNorm kernels derive
                               them from sum/sq (training) or the running buffers (eval) themselves */
  float* coef;              /* backward: (3, Cp) dy = coef0*dz + coef1*y + coef2, written by the last block
                               of mp_bn_bwd_reduce, read by mp_bn_bwd_apply; NULL = recompute per block */
  void* dy;                 /* backward: bf16 (M, Cp) gradient w.r.t. y */
  float* dgamma;            /* backward: (C), accumulated (+=) */
  float* dbeta;             /* backward: (C), accumulated (+=) */
} mp_bn_branch;

typedef struct mp_bn_args {
  mp_bn_branch a, b;
  const void* res;          /* bf16 (M, Cp) identity residual or NULL */
  int32_t relu_a, relu_out;
  void* out;                /* forward output bf16 (M, Cp) (also read by backward when relu_out) */
  float* out_nchw;          /* forward: fp32 (N, C, HW) copy of the output (logits) or NULL */
  const void* dout;         /* backward: bf16 (M, Cp) gradient w.r.t. out, or NULL */
  const float* dout_nchw;   /* backward: fp32 (N, C, HW) gradient w.r.t. out_nchw, or NULL */
  void* dres;               /* backward: bf16 (M, Cp) gradient w.r.t. res, or NULL */
  float* sums;              /* backward workspace (stat_replicas, 4, Cp): zero before mp_bn_bwd_reduce */
  uint32_t* bwd_counter;    /* backward: ticket counter (zero before mp_bn_bwd_reduce) for the coef finalize, or NULL */
  int32_t stat_replicas;    /* >= 1: number of copies of sum / sq / sums the producers spread their atomics over */
  int64_t stat_stride;      /* elements between copies of sum / sq (copies of sums are 4*Cp apart) */
  int64_t M;
  int32_t C, Cp, HW;
  int32_t training;
  float momentum, eps;
  int64_t lo_delta;         /* split (bf16x3) mode: every bf16 (M, Cp) tensor above is a hi / lo pair, lo stored lo_delta
                               elements after hi (see mp_igemm_args.lo_delta); 0 = plain bf16 */
} mp_bn_args;

MP_API int mp_bn_fwd(const mp_bn_args* args, void* stream);
/* backward = two launches: per-channel reductions, then the elementwise gradient. */
MP_API int mp_bn_bwd_reduce(const mp_bn_args* args, void* stream);
MP_API int mp_bn_bwd_apply(const mp_bn_args* args, void* stream);
/* Grouped variants: n_problems (1..MP_MAX_GROUP) BatchNorm passes of identical shape and flags on different
 * tensors in one launch (grid y = problem); args is a contiguous array. */
MP_API int mp_bn_fwd_grouped(const mp_bn_args* args, int n_problems, void* stream);
MP_API int mp_bn_bwd_reduce_grouped(const mp_bn_args* args, int n_problems, void* stream);
MP_API int mp_bn_bwd_apply_grouped(const mp_bn_args* args, int n_problems, void* stream);

/* Deterministic mode (tunable "deterministic", mirrors the reference's `init_algorithms(deterministic=True)`,
 * utils.py:19-24): batch statistics of y = args->a.y as a separate fixed-order pass instead of the conv epilogue's
 * atomics.  Block r sums its share of the pixels and STORES the partial sums into replica r of a.sum / a.sq
 * (stat_replicas copies stat_stride floats apart; mp_bn_fwd adds the replicas in order).  Uses a.y, a.sum, a.sq, M, Cp,
 * stat_replicas, stat_stride, lo_delta. */
MP_API int mp_bn_stats(const mp_bn_args* args, void* stream);
MP_API int mp_bn_stats_grouped(const mp_bn_args* args, int n_problems, void* stream);

/* Inference: eval-mode BatchNorm (running statistics) of n layers as per-channel affines
 *   scale[c] = gamma[c] / sqrt(running_var[c] + eps),  shift[c] = beta[c] - (running_mean[c] - conv_bias[c]) * scale[c]
 * (zeros for C <= c < Cp), which mp_conv_igemm applies in its epilogue (ep_scale / ep_shift) -- BatchNorm folded
 * into the conv, bin/infer_single.py:58-66 / bin/eval_3d.py:60-62.  `table` is a DEVICE array of n entries. */
typedef struct mp_bn_fold_entry {
  const float* gamma;
  const float* beta;
  const float* running_mean;
  const float* running_var;
  const float* conv_bias;   /* or NULL */
  float* scale;             /* (Cp) */
  float* shift;             /* (Cp) */
  int32_t C, Cp;
  float eps;
  int32_t reserved;
} mp_bn_fold_entry;
MP_API int mp_bn_fold_eval(const mp_bn_fold_entry* table, int n, void* stream);

/* `lo_delta` of the entry points below: 0 = plain bf16 activations; > 0 = split (bf16x3) mode, every bf16 activation
 * tensor argument is a hi / lo pair with lo stored lo_delta elements after hi (see mp_igemm_args.lo_delta).  The pure
 * data movements (mp_maxpool_bwd, mp_axis_permute) are linear: call them once per half.
 *
 * nn.MaxPool2d(3, 2, 1) of the ResNet stem on bf16 NHWC; idx (N, H/2, W/2, C) uint8 records the
 * arg-max tap (first maximum in row-major window order, as ATen does) for the backward. */
MP_API int mp_maxpool_fwd(const void* x, void* y, uint8_t* idx, int N, int H, int W, int C, int64_t lo_delta,
                          void* stream);
MP_API int mp_maxpool_bwd(const void* dy, const uint8_t* idx, void* dx, int N, int H, int W, int C,
                          void* stream);

/* HeatmapColumn axis permutation (models/margipose_model.py:86-99) on bf16 NHWC (N, S, S, C),
 * C = G*S real channels inside Cp: mode 1 ('zy') out[n,h,c',g*S+w] = in[n,h,w,g*S+c'];
 * mode 2 ('xz') out[n,c',w,g*S+h] = in[n,h,w,g*S+c'].  Both are involutions (backward = same call). */
MP_API int mp_axis_permute(const void* in, void* out, int mode, int N, int S, int C, int Cp, void* stream);

/* HeatmapCombiner (models/margipose_model.py:142-150) fused with the stage-input update (:195):
 * out[pix, c] = inp[pix, c] + sum_k sum_j w[c, k*J + j] * p_k[n, j, pix]; p_k fp32 (N, J, HW). */
MP_API int mp_combiner_fwd(const float* const p[3], const float* w, const void* inp, void* out,
                           int N, int J, int HW, int C, int64_t lo_delta, void* stream);
/* d p_k (fp32 (N, J, HW); overwritten, or += when accumulate != 0) and d w (+=) from
 * d out (bf16 (N*HW, C)). */
MP_API int mp_combiner_bwd(const void* dout, const float* const p[3], const float* w,
                           float* const dp[3], float* dw, int accumulate, int N, int J, int HW,
                           int C, int64_t lo_delta, void* stream);

/* ResNet stem conv1 (7x7, stride 2, padding 3, 3 input channels): gathers the fp32 NCHW image
 * into bf16 patch rows (N, H/2, W/2, 192) with k = (r*7 + s)*3 + c (147 real + zero padding), so
 * the conv and its weight gradient run as 1x1 cases of mp_conv_igemm / mp_conv_wgrad. */
MP_API int mp_stem_im2col(const float* x, void* patches, int N, int H, int W, int64_t lo_delta, void* stream);
/* Same from a uint8 NHWC image (N, H, W, 3) with the reference's input step fused in: ImageSpecs.convert =
 * to_tensor (/255) + (x - mean) / stddev (data_specs.py:6-13,38-39; mean / stddev: HOST pointers to 3 floats,
 * ImageNet statistics for the MargiPose model, models/margipose_model.py:206-209). */
MP_API int mp_stem_im2col_u8(const uint8_t* x, void* patches, const float mean[3], const float stddev[3],
                             int N, int H, int W, int64_t lo_delta, void* stream);

/* out = sum of n (<= 4) bf16 tensors of `count` elements (gradient fan-in). */
MP_API int mp_add_bf16(const void* const in[4], int n, void* out, int64_t count, int64_t lo_delta, void* stream);

/* fp32 master weights -> bf16 GEMM operands, one launch for the whole network.  `table` is a
 * DEVICE array of n_entries records sorted by work_off; dst elements beyond the source extent are
 * zero-filled.  Entry e fills the [rows_p][taps][cols_p] block
 *   packed[dst_off + r*dst_row_stride + t*cols_p + c] = transpose ? src[c][t][r] : src[r][t][c]
 * (src = master + src_off, fp32 [A][taps][B]); a row stride larger than taps*cols_p lets two layers
 * share one matrix along K (fused data gradient of a residual block).  lo_off > 0 (split mode): the rounding
 * residual bf16(w - bf16(w)) of every element is written lo_off elements after it (the W_lo operand). */
typedef struct mp_pack_entry {
  int64_t src_off;         /* elements into `master` */
  int64_t dst_off;         /* elements into `packed` */
  int64_t dst_row_stride;  /* elements */
  int64_t work_off;        /* cumulative rows_p*taps*cols_p of the entries before this one */
  int64_t work_end;        /* work_off + rows_p*taps*cols_p */
  int32_t A, B, taps;      /* source extents */
  int32_t transpose;
  int32_t rows_p, cols_p;  /* padded destination extents (cols_p % 8 == 0) */
  int64_t lo_off;          /* > 0: also write the rounding residual bf16(w - bf16(w)) lo_off elements after every element
                              (split mode: the matrix is [rows][K_hi | K_lo], lo_off = K of that matrix); 0 = none */
} mp_pack_entry;
MP_API int mp_pack_weights(const float* master, void* packed, const mp_pack_entry* table,
                           int n_entries, int64_t total_work, void* stream);

/* torch.optim.SGD step over flat fp32 buffers (momentum buffer initialised on first_step):
 *   g = grad*grad_scale + wd*p;  buf = first ? g : mom*buf + (1-dampening)*g;
 *   p -= lr * (nesterov ? g + mom*buf : buf) */
MP_API int mp_sgd_step(float* param, const float* grad, float* momentum_buf, int64_t n, float lr,
                       float momentum, float dampening, float weight_decay, int nesterov,
                       int first_step, float grad_scale, void* stream);
/* Same, but when `hyper` (device pointer to 6 floats: lr, momentum, dampening, weight_decay, grad_scale,
 * first_step as 0/1) is given the kernel reads them from it at run time, so a step captured in a CUDA graph
 * follows LR / momentum schedules (the reference's 1-cycle schedule, hyperparam_scheduler.py:24-42;
 * momentum_buf is then required).  nesterov stays a launch-time constant. */
MP_API int mp_sgd_step_hp(float* param, const float* grad, float* momentum_buf, int64_t n, float lr,
                          float momentum, float dampening, float weight_decay, int nesterov,
                          int first_step, float grad_scale, const float* hyper, void* stream);

/* Tunables for experiments (name -> value); returns MP_ERR_ARG for unknown names.
 *   "igemm_smem"  : shared-memory budget per (persistent) CTA of mp_conv_igemm in bytes (default 204800)
 *   "igemm_ctas"  : persistent CTAs per mp_conv_igemm launch (default 0 = the device's SM count)
 *   "igemm_halo"  : 1 (default) = filter taps that differ only by their row shift share one activation box of
 *                   tile_rows + 2 rows (fetched once, read through row-shifted descriptors); 0 = one box per tap
 *   "igemm_pair"  : 1 = two CTAs of a cluster pair up on M=256 tcgen05.mma.cta_group::2 tiles, each loading half of
 *                   every weight tile (when the M tiles pair up); 0 = one CTA per tile (default 1)
 *   "igemm_mt"    : 2 = two 128-pixel row blocks (TMEM accumulators) per tile share every weight tile when the launch
 *                   keeps at least "igemm_mt_ctas" tiles (0 = the CTA count) (default 2); 1 = one
 *   "igemm_split_n": mp_conv_igemm halves its N tile when a launch has fewer tiles than this (default 0 = the CTA count)
 *   "igemm_resident": 1 (default) = a CTA keeps its whole weight operand in shared memory across its tiles when it fits
 *   "igemm_astages": activation (halo box) slots in flight when the weights stream (default 3)
 *   "igemm_prefetch_b": 1 = a CTA requests the first weight tiles of its first work unit before griddepcontrol.wait (the
 *                   packed weights do not depend on the stream predecessor); 0 (default, faster as measured) = after it
 *   "igemm_dbg"   : experiment switches, timing only -- results are garbage (1 = no TMA loads, 2 = no MMA, 4 = no epilogue)
 *   "igemm_trace" : device pointer to 64 uint64: CTA (0,0,0) stamps %globaltimer at its phase boundaries (tools/trace_igemm.py)
 *   "pdl"         : 1 (default) = programmatic dependent launch for the kernels of the network chain
 *   "deterministic": 1 = bit-wise run-to-run reproducible results: mp_conv_wgrad runs without split-K, mp_bn_bwd_reduce
 *                   launches one block per replica of its sums, mp_combiner_bwd adds its weight-gradient partials in block
 *                   order; the caller additionally uses mp_bn_stats instead of mp_conv_igemm's stat_sum / bn arguments
 *                   (the engine does: MargiPoseModel.deterministic / MARGIPOSE_B200_DETERMINISTIC=1).  Default 0.
 *   "wgrad_ctas"  : target CTA count of mp_conv_wgrad (default 148)
 *   "wgrad_halo"  : 1 = up to three row-shifted taps per CTA share the A tile and one halo box of B; 0 (default) = one tap per CTA
 *   "wgrad_slice" : widest column slice of B per CTA when taps are grouped (default 256; 64 or 128 narrow it)
 *   "wgrad_kp"    : pixels per pipeline stage of mp_conv_wgrad (default 128)
 *   "wgrad_smem"  : shared-memory budget of mp_conv_wgrad's pipeline stages in bytes (default 204800)
 *   "wgrad_dbg"   : experiment switches (1 = skip the gradient atomics, 4 = no MMA)
 *   "bn_tma"      : bit mask (1 = mp_bn_fwd, 2 = mp_bn_bwd_reduce, 4 = mp_bn_bwd_apply) of the BatchNorm kernels that stream their
 *                   inputs through shared-memory tile rings filled by cp.async.bulk (TMA) where the tensors are plain bf16
 *                   NHWC; default 1 (the backward kernels share SMs with the weight-gradient kernels, see csrc/bn.cu)
 *   "tail_fast"   : 1 (default) = mp_tail_fwd / mp_tail_bwd use the log2-domain kernels when the heatmap width divides 128;
 *                   0 = the general warp-sliced kernels
 *   "tail_wpj"    : register-resident "warp per joint" kernels (no shared-memory staging) for planes whose row length is a
 *                   power of two <= 128: 0 = off, 1 = planes of at most 1024 elements (the model's 32 x 32 heatmaps; 4096
 *                   without a JS term) in one warp, 2 (default) = also larger planes (multiples of 1024 elements up to
 *                   128 x 128) in 2 ... 16 warps, 3 = same with 8 instead of 16 float4 per lane in the forward (slower)
 *   "tail_wpj_max": most float4 per lane of the one-warp forward kernel without a JS term (default 32 = 64 x 64 planes)
 *   "tail_cap"    : float4 per lane the shared-memory (fast / warp) kernels plan for, 4 (default) or 8
 *   "tail_ctas_per_sm": > 0 caps the grid of the shared-memory (fast) kernels at that many (persistent) CTAs per SM (default 0 = one CTA per (sample, joint) group) */
MP_API int mp_set_tunable(const char* name, int64_t value);

#ifdef __cplusplus
}
#endif
#endif /* MARGIPOSE_B200_H */
