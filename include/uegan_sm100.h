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

#define UEGAN_ABI_VERSION 1

enum { UEGAN_F32 = 0, UEGAN_BF16 = 1 };                 /* storage dtype; F32 tensors feed kind::tf32 MMAs */
enum { UEGAN_ACT_NONE = 0, UEGAN_ACT_LRELU = 1, UEGAN_ACT_RELU = 2, UEGAN_ACT_TANH = 3, UEGAN_ACT_SIGMOID = 4 };
enum { UEGAN_PAD_ZERO = 0, UEGAN_PAD_REFLECT = 1 };

typedef struct uegan_tensor {
  void* data;      /* allocation start = element (n=0, y=-halo, x=-halo, c=0) */
  int32_t n, h, w; /* logical (interior) extent */
  int32_t c;       /* stored channels per pixel (c * sizeof(dtype) must be a multiple of 16) */
  int32_t halo;    /* halo pixels on each side */
  int32_t dtype;   /* UEGAN_F32 | UEGAN_BF16 */
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
  const float* bias;    /* [cout] fp32 or NULL */
  const float* alpha;   /* device scalar (1/sigma of spectral norm, models.py:185-188) or NULL (=1) */
  const uegan_tensor* mul;     /* optional: y *= mul[n,ho,wo,co] after the activation (y4.mul(x1), models.py:70) */
  float* out_nchw;             /* optional planar fp32 output, see above */
  const float* residual_nchw;  /* optional, with out_nchw */
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

int uegan_conv2d_fprop(const uegan_conv_desc* desc, void* stream);

/* NCHW fp32 image batch -> NHWC tensor with halo (pad_mode) and per-channel affine  v = x*scale[c] + shift[c]
 * (scale/shift are HOST arrays of 3 floats or NULL).  Replaces the implicit layout of models.py:47 inputs, and
 * (x+1)/2 -> (x-mean)/std of trainer.py:108 + losses.py:26-27 for the VGG tower.  Channels >= 3 are zero. */
int uegan_pack_input(const float* x_nchw, const uegan_tensor* dst, int32_t pad_mode, const float* scale_host,
                     const float* shift_host, void* stream);
/* Writes the halo of t from its interior (reflect) or with zeros. */
int uegan_halo_fill(const uegan_tensor* t, int32_t pad_mode, void* stream);
/* nn.InstanceNorm2d(affine=False), biased variance, eps (models.py:227,236; losses.py:18):
 * dst[..., dst_c_off + c] = (src[..., c] - mean[n,c]) * rsqrt(var[n,c] + eps).  stats_ws: 2*n*c doubles. */
int uegan_instance_norm(const uegan_tensor* src, const uegan_tensor* dst, int32_t dst_c_off, float eps,
                        double* stats_ws, void* stream);
/* F.interpolate(scale_factor=2, mode='bilinear', align_corners=True) (models.py:191-201) into a channel slice. */
int uegan_upsample2x(const uegan_tensor* src, const uegan_tensor* dst, int32_t dst_c_off, void* stream);
/* nn.MaxPool2d(2,2) of torchvision vgg19.features (losses.py:43). */
int uegan_maxpool2x2(const uegan_tensor* src, const uegan_tensor* dst, void* stream);
/* NHWC tensor interior channels [c_off, c_off+c_count) -> NCHW fp32 (test / debug readback). */
int uegan_unpack_nchw(const uegan_tensor* src, int32_t c_off, int32_t c_count, float* dst_nchw, void* stream);

/* Hardware probe used by tests/DESIGN.md: runs a 128xNx(32*kchunks) tf32 GEMM whose A operand is read from a
 * shared-memory window shifted by `row_shift` 128-byte rows with the given descriptor base_offset; see
 * csrc/probe.cu.  out: 128*n floats. */
int uegan_probe_umma_window(const float* a, const float* b, float* out, int32_t a_rows, int32_t n, int32_t kchunks,
                            int32_t row_shift, int32_t base_offset, int32_t sbo_bytes, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* UEGAN_SM100_H */
