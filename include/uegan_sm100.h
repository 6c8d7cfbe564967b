/* libuegan_sm100.so -- C ABI of the B200-native UEGAN hot path.
 *
 * The reference (eezkni/UEGAN) has no FFI of its own: its hot path is torch.nn calls made from
 * models.py / losses.py / trainer.py.  Each entry point below is what a binding for that path
 * replaces; the reference call site is cited per function.  All pointers are DEVICE pointers
 * unless a name ends in `_host`.  No torch types cross this boundary; the caller owns all memory;
 * every call is asynchronous on `stream` (a cudaStream_t passed as void*) and never synchronises.
 * Return value: 0 on success, negative on error (text via uegan_last_error(), thread-local).
 *
 * Activation tensors are NHWC with an optional halo of `halo` pixels on every side of H and W:
 *   element (n, y, x, c), y in [-halo, h+halo), lives at
 *   data[ ((n*(h+2*halo) + y+halo)*(w+2*halo) + x+halo)*c_total + c ].
 * The halo holds the reflection (nn.ReflectionPad2d, models.py:82,93,161,173) or zero padding
 * (torchvision VGG conv padding=1) the CONSUMER convolution needs, so every convolution is a
 * "valid" convolution over a TMA-addressable padded tensor.
 */
#ifndef UEGAN_SM100_H
#define UEGAN_SM100_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define UEGAN_ABI_VERSION 4

enum { UEGAN_F32 = 0, UEGAN_BF16 = 1, UEGAN_F16 = 2 };  /* storage dtype; F32 tensors feed kind::tf32 MMAs, the
                                                           16-bit types kind::f16 */
enum { UEGAN_ACT_NONE = 0, UEGAN_ACT_LRELU = 1, UEGAN_ACT_RELU = 2, UEGAN_ACT_TANH = 3, UEGAN_ACT_SIGMOID = 4 };
enum { UEGAN_PAD_ZERO = 0, UEGAN_PAD_REFLECT = 1 };

typedef struct uegan_tensor {
  void* data;      /* allocation start = element (n=0, y=-halo, x=-halo, c=0) */
  int32_t n, h, w; /* logical (interior) extent */
  int32_t c;       /* stored channels per pixel (c * sizeof(dtype) must be a multiple of 16) */
  int32_t halo;    /* halo pixels on each side */
  int32_t dtype;   /* UEGAN_F32 | UEGAN_BF16 | UEGAN_F16 */
  const float* scale; /* NULL, or a DEVICE scalar s > 0 (a power of two): the stored values are s * the true values.
                         fp16 tensors of the Generator / Discriminator carry one (their magnitudes follow the weights: 3e-9
                         after five layers of the reference's orthogonal(0.02) init, far below fp16's range); every kernel
                         divides it out of what it loads and multiplies its output by the destination's scale, exactly
                         (powers of two), so results are those of the unscaled arithmetic.  uegan_scale_update maintains it. */
} uegan_tensor;

/* One implicit-GEMM convolution (tcgen05.mma + TMA), forward.
 * Replaces nn.ReflectionPad2d + nn.Conv2d (+ LeakyReLU / Tanh / clamp(res + x)) at models.py:82-83,93-97,
 * 161-165,173-178,70-72 and torchvision vgg19 Conv2d+ReLU (losses.py:43-116).
 *   y[n, ho, wo, y_c_off + co] = epi( alpha * sum_{r,s,ci} x[n, ho*stride + r - pad, wo*stride + s - pad, ci]
 *                                               * w[co, ci, r, s] + bias[co] )
 * x.halo >= pad is required (the halo must already hold the padding values).
 * If out_nchw != NULL (cout <= 16) the result is written as fp32 NCHW planes instead of y:
 *   out_nchw[n][co][ho][wo] = residual_nchw ? clamp(act(.) + residual_nchw[n][co][ho][wo], -1, 1) : act(.)   */
typedef struct uegan_conv_desc {
  uegan_tensor x;
  uegan_tensor y;
  int32_t y_c_off;      /* first channel of y written by this conv (concat-by-construction, models.py:55-67) */
  int32_t cout;         /* output channels produced */
  int32_t k;            /* square kernel size */
  int32_t stride;       /* 1 or 2 */
  int32_t pad;          /* (k-1)/2 in the reference, must be <= x.halo */
  int32_t act;          /* UEGAN_ACT_* */
  const void* w_packed; /* from uegan_pack_conv_weight, same dtype as x */
  const float* w_scale; /* NULL, or the device scalar the packed weight was multiplied by (uegan_pack_conv_weight_scaled) */
  const float* bias;    /* [cout] fp32 or NULL */
  const float* alpha;   /* device scalar (1/sigma of spectral norm, models.py:185-188) or NULL (=1) */
  const uegan_tensor* mul;     /* optional: y *= mul[n,ho,wo,co] after the activation (y4.mul(x1), models.py:70) */
  float* out_nchw;             /* optional planar fp32 output, see above */
  const float* residual_nchw;  /* optional, with out_nchw */
  float* aux_nchw;             /* optional, with out_nchw: act(.) before the residual / clamp (saved for backward) */
  int32_t y_mul;               /* 0/1: dense output.  2: output (a, b) is written at (y_off_h + 2a, y_off_w + 2b) of y
                                  (one parity class of the dgrad of a stride-2 conv per launch) */
  int32_t y_off_h, y_off_w;
  const uegan_tensor* mask;    /* optional: y *= act'(mask[n,ho,wo,co]) (LRELU: 1 / 0.2, RELU: 1 / 0 by the sign of the
                                  forward activation) -- fuses the activation backward into a dgrad launch */
  int32_t mask_act;
  double* in_stats;            /* optional [n][cout][2]: the epilogue accumulates sum / sum of squares of the stored
                                  outputs per (n, c) (zeroed by the call); consumed by uegan_instance_norm_apply.
                                  Needs Ho*Wo >= 128 (tiles within one image). */
  const uegan_tensor* y_premul; /* optional, with `mul` (16-bit tensors): the output BEFORE the multiplication is stored here
                                  as well (own scale) -- training keeps y4 next to y4.mul(x1), models.py:70 */
  int32_t y_reflect_halo;      /* != 0: the epilogue also writes y's halo as the reflection padding of its interior
                                  (nn.ReflectionPad2d of the consumer, models.py:82,93,161,173): uegan_halo_fill(y, REFLECT)
                                  without a separate pass.  16-bit dense outputs, y.h, y.w > 2 * y.halo + 1 */
  int32_t y_cls_c;             /* 0, or (with y_mul = 2, y_off = 0) ALL FOUR parity classes of a stride-2 data gradient in
                                  one launch: cout = 4 * y_cls_c output columns, column (pi*2 + pj) * y_cls_c + c is
                                  channel c of the output pixel (2a + pi, 2b + pj); w_packed = the four class operands
                                  of uegan_pack_conv_weight_dgrad back to back (class = pi*2 + pj).  Small-Cin layers
                                  get 4x wider MMAs, deep ones one grid instead of four quarter-filled ones. */
} uegan_conv_desc;

int uegan_abi_version(void);
const char* uegan_last_error(void);
/* Reads and clears the device-side watchdog flag (non-zero = a bounded mbarrier wait expired). Synchronises. */
int uegan_device_error(void);

/* Bytes needed for the packed weight of a conv with `cout` x `cin` x k x k weights whose input tensor stores
 * `cin_stored` channels (>= cin, channel padding) in `dtype`. */
size_t uegan_packed_weight_bytes(int32_t cout, int32_t cin_stored, int32_t k, int32_t dtype);
/* OIHW fp32 (nn.Conv2d.weight, models.py:83,94,162,174) -> K-major packed [cout_pad][k][row_pad] of `dtype`
 * (tf32-rounded for UEGAN_F32).  `cin_first`/`cin` select input channels [cin_first, cin_first+cin) of a weight
 * with `cin_total` input channels (GAM fuse uses the first half only, models.py:225,234).
 * transpose_flip != 0 packs the dgrad operand instead: roles of O and I swapped, taps rotated by 180 degrees. */
int uegan_pack_conv_weight(const float* w_oihw, void* w_packed, int32_t cout, int32_t cin_total, int32_t cin_first,
                           int32_t cin, int32_t cin_stored, int32_t k, int32_t dtype, int32_t transpose_flip,
                           void* stream);

/* dgrad operand of a conv with weight w (cout_orig x cin_total x k_orig x k_orig, stride 1 or 2): the data gradient
 * is itself a stride-1 valid convolution of the zero-haloed output gradient dz (cout_stored channels per pixel) with
 * this operand; for stride 2 one operand per parity class (pi, pj) of the input position, kernel size ceil(k/2).
 * Packed size: uegan_packed_weight_bytes(cin, cout_stored, ceil(k_orig/stride), dtype). */
int uegan_pack_conv_weight_dgrad(const float* w_oihw, void* w_packed, int32_t cout_orig, int32_t cin_total,
                                 int32_t cin_first, int32_t cin, int32_t cout_stored, int32_t k_orig, int32_t stride,
                                 int32_t pi, int32_t pj, int32_t dtype, void* stream);
/* Both packers with a scale: the packed values are w * (*w_scale_dev) (a device scalar maintained by
 * uegan_scale_update on the fp32 master weight); pass the same pointer as conv_desc.w_scale. */
int uegan_pack_conv_weight_scaled(const float* w_oihw, void* w_packed, int32_t cout, int32_t cin_total, int32_t cin_first,
                                  int32_t cin, int32_t cin_stored, int32_t k, int32_t dtype, const float* w_scale_dev,
                                  void* stream);
int uegan_pack_conv_weight_dgrad_scaled(const float* w_oihw, void* w_packed, int32_t cout_orig, int32_t cin_total,
                                        int32_t cin_first, int32_t cin, int32_t cout_stored, int32_t k_orig,
                                        int32_t stride, int32_t pi, int32_t pj, int32_t dtype, const float* w_scale_dev,
                                        void* stream);
/* The four parity-class operands (pi, pj) = (0,0), (0,1), (1,0), (1,1) of a STRIDE-2 conv's data gradient, back to back
 * (class stride = uegan_packed_weight_bytes(cin, cout_stored, ceil(k_orig / 2), dtype)), in one launch: the operand of a
 * uegan_conv2d_fprop call with y_cls_c.  w_scale_dev may be NULL. */
int uegan_pack_conv_weight_dgrad4(const float* w_oihw, void* w_packed, int32_t cout_orig, int32_t cin_total, int32_t cin_first,
                                  int32_t cin, int32_t cout_stored, int32_t k_orig, int32_t dtype, const float* w_scale_dev,
                                  void* stream);

/* Per-tensor power-of-two scales (uegan_tensor.scale), maintained on the device so that CUDA-graph replays need no host
 * work.  One entry per tensor: the kernel samples up to `max_samples` stored values evenly across the buffer, takes
 * their largest magnitude, and sets  *scale = 2^floor(log2(target / (amax_stored / *scale)))  -- the scale the NEXT
 * pass uses (delayed scaling: magnitudes drift slowly, `target` = 2^10 leaves a 64x margin below fp16's 65504).  A sample
 * that overflowed (inf / nan) divides the scale by 2^8; an all-zero 16-bit sample multiplies it by 2^8 (underflow: the next
 * pass measures again), an all-zero fp32 sample leaves it unchanged.  dtype F32 entries are
 * fp32 buffers (master weights: their scale is what the packers multiply by). */
typedef struct uegan_scale_entry {
  const void* data;   /* device buffer */
  int64_t numel;      /* elements in the buffer */
  int32_t dtype;      /* UEGAN_F32 | UEGAN_F16 | UEGAN_BF16 */
  int32_t reserved;   /* 0: the buffer holds stored (= scale * true) values; 1: it holds TRUE values (fp32 master weights whose
                         scale is what uegan_pack_conv_weight*_scaled multiplies by) */
  float* scale;       /* device scalar to update */
} uegan_scale_entry;
int uegan_scale_update(const uegan_scale_entry* entries_dev, int32_t count, float target, int32_t max_samples, void* stream);

int uegan_conv2d_fprop(const uegan_conv_desc* desc, void* stream);

/* "Row-sum" variant of uegan_conv2d_fprop for stride-1 convolutions with a TINY output-channel count written as planar
 * fp32 (desc->out_nchw): the Generator's last conv 32 -> 3, k7 + tanh + clamp(res + x) (models.py:34-35,72) and the
 * Discriminator's prediction heads C -> 1, k7 / k5 (models.py:110-126).  The horizontal taps become GEMM columns
 * (N = k*cout), only the vertical taps stay in the K loop, and the epilogue sums the k shifted columns with warp
 * shuffles: k*C/8 MMAs per 128 PATCH pixels instead of k*k*C/8 per 128 output pixels.  tf32 only, x.c % 32 == 0,
 * (k, cout) in {(7,3), (7,1), (5,1), (3,1)}.  desc->w_packed must come from uegan_pack_conv_weight_rowsum; y, mul, mask,
 * in_stats, y_mul are not used.  uegan_conv2d_rowsum_supported() != 0 tells whether a shape qualifies (weights plus two
 * patch stages must fit in shared memory). */
int uegan_conv2d_rowsum_supported(int32_t cout, int32_t cin_stored, int32_t k, int32_t dtype);
size_t uegan_packed_weight_rowsum_bytes(int32_t cout, int32_t cin_stored, int32_t k);
/* OIHW fp32 -> [(s, o) rows padded to a multiple of 16][r][cin_stored] tf32. */
int uegan_pack_conv_weight_rowsum(const float* w_oihw, void* w_packed, int32_t cout, int32_t cin_total, int32_t cin_first,
                                  int32_t cin, int32_t cin_stored, int32_t k, void* stream);
/* The same operand multiplied by a device scalar and stored as `dtype` (UEGAN_F32: tf32-rounded; UEGAN_F16: rows padded to
 * whole 64-channel chunks) -- the fp16 row-sum path; pass the scalar as conv_desc.w_scale. */
int uegan_pack_conv_weight_rowsum_scaled(const float* w_oihw, void* w_packed, int32_t cout, int32_t cin_total,
                                         int32_t cin_first, int32_t cin, int32_t cin_stored, int32_t k, int32_t dtype,
                                         const float* w_scale_dev, void* stream);
int uegan_conv2d_fprop_rowsum(const uegan_conv_desc* desc, void* stream);

/* NCHW fp32 image batch -> NHWC tensor with halo (pad_mode) and per-channel affine  v = x*scale[c] + shift[c]
 * (scale/shift are HOST arrays of 3 floats or NULL).  Replaces the implicit layout of models.py:47 inputs, and
 * (x+1)/2 -> (x-mean)/std of trainer.py:108 + losses.py:26-27 for the VGG tower.  Channels >= 3 are zero. */
int uegan_pack_input(const float* x_nchw, const uegan_tensor* dst, int32_t pad_mode, const float* scale_host,
                     const float* shift_host, void* stream);
/* Writes the halo of t from its interior (reflect) or with zeros. */
int uegan_halo_fill(const uegan_tensor* t, int32_t pad_mode, void* stream);
/* nn.InstanceNorm2d(affine=False), biased variance, eps (models.py:227,236; losses.py:18):
 * dst[..., dst_c_off + c] = (src[..., c] - mean[n,c]) * rsqrt(var[n,c] + eps).  stats_ws: 3*n*c doubles. */
int uegan_instance_norm(const uegan_tensor* src, const uegan_tensor* dst, int32_t dst_c_off, float eps,
                        double* stats_ws, void* stream);
/* Second half of uegan_instance_norm for statistics produced by a convolution epilogue (conv_desc.in_stats):
 * stats_ws holds [n][c][2] sums followed by room for n*c (mean, rstd) float pairs (3*n*c doubles in total). */
int uegan_instance_norm_apply(const uegan_tensor* src, const uegan_tensor* dst, int32_t dst_c_off, float eps,
                              double* stats_ws, void* stream);
/* F.interpolate(scale_factor=2, mode='bilinear', align_corners=True) (models.py:191-201) into a channel slice. */
int uegan_upsample2x(const uegan_tensor* src, const uegan_tensor* dst, int32_t dst_c_off, void* stream);
/* One decoder concat of the Generator in a single pass (models.py:55-67 `torch.cat([upsample(y), GAM(skip)], 1)`):
 * dst[.., 0:C) = bilinear x2 (align_corners=True) of u, dst[.., C:2C) = (z - mean) * rstd with the finalised pairs of
 * uegan_instance_norm_stats(z).  Whole 2C-channel pixels are written (the two separate passes wrote half lines). */
int uegan_cat_build(const uegan_tensor* u, const uegan_tensor* z, const float* mean_rstd, const uegan_tensor* dst,
                    void* stream);
/* nn.MaxPool2d(2,2) of torchvision vgg19.features (losses.py:43). */
int uegan_maxpool2x2(const uegan_tensor* src, const uegan_tensor* dst, void* stream);
/* NHWC tensor interior channels [c_off, c_off+c_count) -> NCHW fp32 (test / debug readback). */
int uegan_unpack_nchw(const uegan_tensor* src, int32_t c_off, int32_t c_count, float* dst_nchw, void* stream);

/* Spectral norm of one conv weight viewed as rows x cols (torch.nn.utils.spectral_norm, models.py:185-188):
 * train != 0: v <- normalize(W^T u), u <- normalize(W v) in place; sigma_out[0] = u.(W v), sigma_out[1] = 1/sigma.
 * ws: rows + cols + 8 floats of scratch. */
int uegan_spectral_sigma(const float* w, float* u, float* v, int32_t rows, int32_t cols, int32_t train,
                         float* sigma_out, float* ws, void* stream);
/* The same for `count` (<= 8) independent layers at once -- one launch per phase for the whole Discriminator (models.py:139-
 * 155: five spectrally-normalised convs per pass) instead of four per layer.  Tables are HOST arrays of device pointers;
 * ws[l] as above.  u_used / v_used (optional, may hold NULLs): copies of the u / v the pass ends with, which the backward
 * of THIS pass needs after later passes have moved u / v on. */
int uegan_spectral_sigma_batch(int32_t count, const float* const* w, float* const* u, float* const* v, const int32_t* rows,
                               const int32_t* cols, int32_t train, float* const* sigma_out, float* const* ws,
                               float* const* u_used, float* const* v_used, void* stream);

/* Relativistic average GAN loss summed over `nscales` prediction maps (GANLoss.__call__, losses.py:393-409 with
 * gan_mode 'rahinge' (mode 0, losses.py:348-362) or 'rals' (mode 1, :363-377)).  real[i] / fake[i]: fp32 maps of
 * counts[i] elements (the mean is over the whole map incl. batch).  ws: 48 doubles, kept for the backward call.
 * HOST arrays of device pointers. */
int uegan_gan_loss_fwd(int32_t mode, int32_t for_discriminator, int32_t nscales, const float* const* real,
                       const float* const* fake, const int64_t* counts, double* ws, float* loss_out, void* stream);
/* The same loss in three phases for data-parallel training (the relativistic means are over the GLOBAL
 * batch): phase 0 local sums -> ws[0..2*nscales) ; [all-reduce ws[0..16)] ; phase 1 hinge terms and their derivative
 * sums -> ws[16..48) ; [all-reduce ws[16..48)] ; phase 2 loss_out = global loss.  world = number of ranks. */
int uegan_gan_loss_phase(int32_t phase, int32_t mode, int32_t for_discriminator, int32_t nscales,
                         const float* const* real, const float* const* fake, const int64_t* counts, int32_t world,
                         double* ws, float* loss_out, void* stream);
/* d_real[i] / d_fake[i] (either may be NULL) += gscale * dloss/dmap, gscale = gscale_host * (gscale_dev ? *gscale_dev : 1).
 * Includes the path through the batch-global means.  Data parallel: pass gscale_host = -(world size). */
int uegan_gan_loss_bwd(int32_t mode, int32_t for_discriminator, int32_t nscales, const float* const* real,
                       const float* const* fake, const int64_t* counts, const double* ws, float* const* d_real,
                       float* const* d_fake, const float* gscale_dev, float gscale_host, void* stream);
/* One PerceptualLoss term (losses.py:30-34): loss_inout[0] += weight * mean( (IN(x) - IN(y))^2 ), with the per-(n,c)
 * (mean, rstd) pairs of x and y as produced by uegan_instance_norm_stats / a conv epilogue.  accum: 1 double (zero). */
int uegan_in_mse_fwd(const uegan_tensor* x, const uegan_tensor* y, const float* mean_rstd_x, const float* mean_rstd_y,
                     float weight, double* accum, float* loss_inout, void* stream);
/* MultiscaleRecLoss.forward (losses.py:219-231) on fp32 NCHW images: type 0 l1, 1 smoothl1, 2 l2; scales 1..3 with
 * weights 1, 1/2, 1/4 and AvgPool2d(2) between scales.  If grad_nchw != NULL also writes
 * grad_scale * (gscale_dev ? *gscale_dev : 1) * dloss/dpred. */
int uegan_msrec_loss(const float* pred_nchw, const float* gt_nchw, int32_t n, int32_t c, int32_t h, int32_t w,
                     int32_t type, int32_t scales, double* accum, float* loss_out, float* grad_nchw, float grad_scale,
                     const float* gscale_dev, void* stream);
/* Per-(n,c) InstanceNorm statistics only: stats_ws gets [n][c][2] sums (doubles) followed by n*c (mean, rstd) float
 * pairs; with sums_ready != 0 the sums are taken as already accumulated (conv_desc.in_stats) and only finalised.
 * Returns the device pointer of the (mean, rstd) pairs in *mean_rstd_out. */
int uegan_instance_norm_stats(const uegan_tensor* src, float eps, double* stats_ws, int32_t sums_ready,
                              float** mean_rstd_out, void* stream);

/* ---- backward pass ---------------------------------------------------------------------------------------------
 * Weight gradient (autograd of nn.Conv2d, `aten::convolution_backward`): dw_oihw[o][cin_first + c][r][s] += scale *
 * (alpha ? *alpha : 1) * sum_pix dz[pix][o] * xpad[pix*stride + (r,s)][c].  fp32 NHWC tensors (kind::tf32, MN-major
 * operands), dz must store a multiple of 32 channels, x 4 or a multiple of 32.  The caller zeroes dw_oihw. */
int uegan_conv2d_wgrad(const uegan_tensor* x, const uegan_tensor* dz, int32_t cout, int32_t cin, int32_t cin_total,
                       int32_t cin_first, int32_t k, int32_t stride, int32_t pad, float* dw_oihw,
                       const float* alpha_dev, float scale, float* ws, size_t ws_bytes, void* stream);
/* (ws, ws_bytes): optional split-K workspace.  NULL: the k-slices publish with fp32 atomics (fastest; the summation order
 * and hence the last bits vary from run to run).  Non-NULL: every slice stores its partial plane into ws and a second
 * kernel adds the planes to dw_oihw in slice order -- bit-reproducible like the reference under
 * cudnn.deterministic=True (utils.py:154); the k-split is clamped to what ws_bytes holds. */
/* The same weight gradient for a stride-1 conv with a tiny output-channel count (G's last conv, D's prediction heads) from
 * the stacked gradient e = uegan_dz_hstack(dz):  dW[o][c][r][s] += scale * alpha * sum_{y,q} xpad[y + r][q][c] * e[y][q][(s,o)]
 * -- the wgrad of a k x 1 convolution with k*cout output channels: one accumulator per (32-channel chunk, 4 filter rows)
 * instead of one per (chunk, row, 4 columns).  pad must be (k-1)/2 <= x.halo; k*cout <= 32. */
int uegan_conv2d_wgrad_hstack(const uegan_tensor* x, const uegan_tensor* e, int32_t cout, int32_t cin, int32_t cin_total,
                              int32_t cin_first, int32_t k, int32_t pad, float* dw_oihw, const float* alpha_dev,
                              float scale, float* ws, size_t ws_bytes, void* stream);
/* The same stride-1 weight gradient with the stack read as a SLIDING WINDOW over dz itself (no uegan_dz_hstack pass, no
 * stack tensor): dz (fp16) must carry a ZERO halo of >= k - 1 pixels; the k * dz.c contiguous values starting k - 1 pixels
 * left of a pixel are its stack row.  Any Cout <= dz.c with k * dz.c <= 256 columns (tiny heads: dz.c = 8; the 32- and
 * 64-channel full-resolution decoder layers of G with k = 3), x.c a multiple of 32, k odd, pad = (k-1)/2.
 * Replaces autograd's aten::convolution_backward (weight half) for models.py:94 (dec3, dec4, dec5) and :174 (heads). */
int uegan_conv2d_wgrad_zwin_supported(int32_t cout, int32_t dz_c, int32_t dz_halo, int32_t x_c, int32_t k, int32_t stride,
                                      int32_t dtype);
int uegan_conv2d_wgrad_zwin(const uegan_tensor* x, const uegan_tensor* dz, int32_t cout, int32_t cin, int32_t cin_total,
                            int32_t cin_first, int32_t k, int32_t pad, float* dw_oihw, const float* alpha_dev, float scale,
                            float* ws, size_t ws_bytes, void* stream);
/* Kernels launched by the calling thread's last uegan_conv2d_wgrad* call: 1 (k-split of one, or atomics) or 2 (GEMM +
 * ordered reduction of the partial planes).  Launch accounting only. */
int uegan_wgrad_last_launches(void);
/* Gradient of the planar heads into a zero-haloed NHWC tensor: mode 0 tanh (models.py:178), 1 sigmoid, 2
 * clamp(tanh(z) + x, -1, 1) with out_nchw = tanh(z) (models.py:35,72). */
int uegan_head_bwd(const float* dout_nchw, const float* out_nchw, const float* x_nchw, int32_t channels, int32_t mode,
                   const uegan_tensor* dz, void* stream);
/* dst[.., dst_c_off + c] = act'(mask) * mul * ( fold(src_a) + add_b + add_c ) for c < channels; halo of dst := 0.
 * src_a has extent (h + 2*pad_a, w + 2*pad_a): the gradient w.r.t. a PADDED conv input; reflection padding folds
 * its border back (adjoint of nn.ReflectionPad2d), zero padding crops.  Any of src_a / add_b / add_c may be NULL. */
int uegan_grad_combine(const uegan_tensor* dst, int32_t dst_c_off, int32_t channels, const uegan_tensor* src_a,
                       int32_t a_c_off, int32_t pad_a, int32_t pad_mode_a, const uegan_tensor* add_b, int32_t b_c_off,
                       const uegan_tensor* add_c, int32_t c_c_off, const uegan_tensor* mask, int32_t mask_c_off,
                       int32_t act, const uegan_tensor* mul, int32_t mul_c_off, void* stream);
/* The same with a second output from the same pass: dst2 (halo 0) = mul2 * ( fold(src_a) + add_b + add_c ), i.e. the sum
 * before mask / mul.  Both factors of y4.mul(x1) (models.py:70) get their gradients from ONE read of the incoming one. */
int uegan_grad_combine2(const uegan_tensor* dst, int32_t dst_c_off, int32_t channels, const uegan_tensor* src_a,
                        int32_t a_c_off, int32_t pad_a, int32_t pad_mode_a, const uegan_tensor* add_b, int32_t b_c_off,
                        const uegan_tensor* add_c, int32_t c_c_off, const uegan_tensor* mask, int32_t mask_c_off,
                        int32_t act, const uegan_tensor* mul, int32_t mul_c_off, const uegan_tensor* dst2,
                        int32_t dst2_c_off, const uegan_tensor* mul2, int32_t mul2_c_off, void* stream);
/* Adjoint of nn.ReflectionPad2d IN PLACE: t (halo = pad) holds the gradient w.r.t. the reflect-padded input of a conv
 * (what the dgrad launches of uegan_conv2d_fprop write, extent (h + 2 pad) x (w + 2 pad)); interior pixels within `pad`
 * of an edge receive their reflected halo copies, then the halo is zeroed.  Same result as uegan_grad_combine with src_a
 * only, but touches O(perimeter) data; the tensor is afterwards a zero-haloed dgrad / wgrad operand. */
int uegan_fold_inplace(const uegan_tensor* t, void* stream);
/* Horizontally unrolled gradient of a tiny-Cout stride-1 conv (cout 1 or 3, k <= 7): dz = 4-channel fp32 NHWC with a zero
 * halo >= k - 1 (uegan_head_bwd); e = 32-channel fp32 NHWC, halo 0, extent h x (w + k - 1):
 * e[n, y, q, s*cout + o] = dz[n, y, q - s, o].  Operand of uegan_conv2d_wgrad_hstack. */
int uegan_dz_hstack(const uegan_tensor* dz, int32_t cout, int32_t k, const uegan_tensor* e, void* stream);
/* out[c] = sum over n, h, w of src[.., c_off + c]  (bias gradient). */
int uegan_channel_sum(const uegan_tensor* src, int32_t c_off, int32_t channels, float* out, int32_t out_channels,
                      int32_t accumulate, void* stream);
/* InstanceNorm backward: dz = rstd * (dout - mean(dout) - xhat * mean(dout * xhat)); ws: 2*n*c doubles. */
int uegan_instance_norm_bwd(const uegan_tensor* dout, int32_t d_c_off, const uegan_tensor* z, const float* mean_rstd,
                            const uegan_tensor* dz, double* ws, void* stream);
/* Adjoint of uegan_upsample2x: dsrc = up^T(dout[.., d_c_off : d_c_off + dsrc.c]). */
int uegan_upsample2x_bwd(const uegan_tensor* dout, int32_t d_c_off, const uegan_tensor* dsrc, void* stream);
/* MaxPool2d(2,2) backward fused with the ReLU mask of the pooled activation (fp16 activations and gradients). */
int uegan_maxpool2x2_bwd(const uegan_tensor* src, const uegan_tensor* dpool, const uegan_tensor* dsrc, void* stream);
/* Backward of one PerceptualLoss term w.r.t. the fp16 feature map x (+ optional gradient from deeper layers, then
 * the tap's own ReLU mask): dx (fp16, zero halo).  Upstream scale = weight * (gscale_dev ? *gscale_dev : 1); the
 * caller folds a loss scale into `weight` (and divides it out in uegan_unpack_input_grad) to stay inside fp16's range. */
int uegan_in_mse_bwd(const uegan_tensor* x, const uegan_tensor* y, const float* mean_rstd_x, const float* mean_rstd_y,
                     float weight, const float* gscale_dev, const uegan_tensor* deep, const uegan_tensor* dx, double* ws,
                     void* stream);
/* One PerceptualLoss term in ONE pass over the two feature maps (losses.py:22-36: InstanceNorm2d of both taps + MSE): the
 * five raw moments per (n, c) -- sum x, x^2, y, y^2, xy -- give the (mean, rstd) pairs of both maps, the loss term
 * (loss_inout += weight * mean((IN(x) - IN(y))^2); accum: one zeroed double of scratch) and the two per-(n, c) sums the
 * backward pass needs, all in closed form.  ws: 9 * n * c doubles; *mrx_out / *mry_out / *sums_out point into it and feed
 * uegan_in_mse_bwd_apply.  eps is the InstanceNorm eps in units of the stored values (times x.scale^2 if x carries one). */
int uegan_in_mse_joint(const uegan_tensor* x, const uegan_tensor* y, float eps, float weight, double* ws, double* accum,
                       float* loss_inout, float** mrx_out, float** mry_out, double** sums_out, void* stream);
/* uegan_in_mse_bwd without its statistics pass: `sums` = the per-(n, c) {sum e, sum e * xhat} from uegan_in_mse_joint. */
int uegan_in_mse_bwd_apply(const uegan_tensor* x, const uegan_tensor* y, const float* mean_rstd_x, const float* mean_rstd_y,
                           float weight, const float* gscale_dev, const uegan_tensor* deep, const uegan_tensor* dx,
                           const double* sums, void* stream);
/* Gradient of uegan_pack_input: NHWC (first 3 channels) -> NCHW fp32 times scale_host[c].  With skip_dout_nchw != NULL the
 * Generator's identity path is added: out = clamp(res + x, -1, 1) (models.py:72) passes skip_dout where |res + x| <= 1. */
int uegan_unpack_input_grad(const uegan_tensor* dx, const float* scale_host, float* dst_nchw, const float* skip_dout_nchw,
                            const float* skip_res_nchw, const float* skip_x_nchw, void* stream);

/* Backward of spectral normalisation through sigma (u, v constant, torch spectral_norm semantics): in place on
 * grad_inout = (dL/dW_sn)/sigma (rows x cols): grad -= (<grad, W>/sigma) u v^T.  sigma = {sigma, 1/sigma}; ws: 1 double. */
int uegan_spectral_bwd(float* grad_inout, const float* w, const float* u, const float* v, const float* sigma,
                       int32_t rows, int32_t cols, double* ws, float* accum_out, void* stream);

/* ---- optimizer + data-parallel gradient reduction (trainer.py:337-338 torch.optim.Adam) ------------------
 * Adam over ONE flat fp32 bucket (all parameters of a network; gradients, exp_avg, exp_avg_sq laid out alike), torch
 * semantics: g += weight_decay * p; exp_avg.lerp_(g, 1 - beta1); exp_avg_sq = beta2 * exp_avg_sq + (1 - beta2) g^2;
 * p -= lr / (1 - beta1^t) * exp_avg / (sqrt(exp_avg_sq) / sqrt(1 - beta2^t) + eps).  state3 = {t, lr / bc1, 1 / sqrt(bc2)}
 * (device floats; t is incremented by the call), lr_dev = device scalar: a CUDA-graph replay needs no host work. */
int uegan_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float* state3,
                    const float* lr_dev, float beta1, float beta2, float eps, float weight_decay, void* stream);
/* The same update with the gradient all-reduce FUSED into it: peer_grads_host[r] = device address of rank r's gradient
 * bucket in this process's address space (peer memory over NVLink / NVSwitch, e.g. torch symmetric memory); the kernel
 * sums the `world` buckets in rank order as it loads them, so every rank applies bit-identical updates and no reduced
 * gradient is ever written.  The caller orders it after all ranks' backward passes (a device-side barrier) and keeps the
 * buckets untouched until all ranks have read them. */
int uegan_adam_step_peers(float* param, const float* const* peer_grads_host, int32_t world, float* exp_avg,
                          float* exp_avg_sq, int64_t n, float* state3, const float* lr_dev, float beta1, float beta2,
                          float eps, float weight_decay, void* stream);
/* out[i] = sum_r peers_host[r][i], i < count <= 64, in rank order: the batch-global sums of the relativistic GAN loss
 * (losses.py:351-360) over peer memory instead of a NCCL all-reduce. */
int uegan_peer_sum_f64(double* out, const double* const* peers_host, int32_t world, int32_t count, void* stream);
/* cudaMemsetAsync(ptr, 0, bytes) on `stream`: zeroing of gradient buckets / per-pass scratch (torch zero_() replacement). */
int uegan_memset_zero(void* ptr, size_t bytes, void* stream);
/* cudaStreamIsCapturing(stream): 0 = not capturing, 1 = capture active, 2 = capture invalidated (debug aid: which call broke
 * a CUDA-graph capture), -1 = query failed. */
int uegan_capture_status(void* stream);

/* ---- the conversions either side of the hot path -------------------------------------------------
 * uint8 HWC RGB batch (n x h x w x 3, device memory) -> what transforms.ToTensor() + Normalize(mean, std) produce
 * (data_loader.py:79-81,100-103: mean = std = 0.5; losses.py:26-27: ImageNet constants), bit for bit:
 *   v = ((float)u8 / 255 - mean[c]) / std[c]
 * written (a) into dst, an NHWC activation tensor whose halo is filled with pad_mode (the Generator / Discriminator / VGG
 * input operand; channels >= 3 are zero; dst may be NULL) and (b) as fp32 NCHW planes x_nchw_out (the tensor the
 * reference's loader yields: residual of models.py:72, target of the losses; may be NULL). */
int uegan_pack_input_u8(const uint8_t* img_nhwc_u8, int32_t n, int32_t h, int32_t w, const uegan_tensor* dst,
                        float* x_nchw_out, int32_t pad_mode, const float* mean_host, const float* std_host, void* stream);
/* fp32 NCHW in [-1, 1] -> uint8 HWC as the Tester writes it: denorm (utils.py:128-130) = clamp((x + 1) / 2, 0, 1), then
 * torchvision.utils.save_image's quantisation clamp(v * 255 + 0.5, 0, 255) -> uint8 (tester.py:70-75), bit for bit. */
int uegan_unpack_output_u8(const float* x_nchw, uint8_t* out_nhwc_u8, int32_t n, int32_t c, int32_t h, int32_t w,
                           void* stream);

/* ---- validation metrics on uint8 HWC image pairs ----------------------------------------------------
 * Sum of squared differences over the image minus a `crop`-pixel border, per image, as an exact integer
 * (metrics/CalcPSNR.py:47-52,85-92: PSNR = 10 log10(255^2 / (sse / count)), count = (h-2crop)(w-2crop)c). */
int uegan_sse_u8(const uint8_t* a_nhwc, const uint8_t* b_nhwc, int32_t n, int32_t h, int32_t w, int32_t c, int32_t crop,
                 uint64_t* sse_out, void* stream);
/* Sum over channels and valid pixels of the SSIM map of skimage.metrics.structural_similarity(multichannel=True,
 * data_range=255) -- 7x7 uniform window, sample covariance, K1 = 0.01, K2 = 0.03 -- evaluated on the image minus a
 * `crop`-pixel border (metrics/CalcSSIM.py:47-62); mssim = sum / ((h-2crop-6)(w-2crop-6)c).  fp64. */
int uegan_ssim_u8(const uint8_t* a_nhwc, const uint8_t* b_nhwc, int32_t n, int32_t h, int32_t w, int32_t c, int32_t crop,
                  double* ssim_sum_out, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UEGAN_SM100_H */
