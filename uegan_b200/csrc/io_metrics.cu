// the steps either side of the hot path, on the GPU.
//   N3  uint8 HWC image -> normalised NHWC (+ halo) operand and fp32 NCHW planes  (transforms.ToTensor + Normalize,
//       data_loader.py:79-81,100-103; the ImageNet variant of losses.py:26-27 is the same kernel with other constants)
//       fp32 NCHW generator output -> uint8 HWC  (denorm, utils.py:128-130, + torchvision save_image's quantisation)
//   N4  PSNR (metrics/CalcPSNR.py:85-92, 4-pixel border crop of :47-52) and SSIM (metrics/CalcSSIM.py:62:
//       skimage.metrics.structural_similarity, 7x7 uniform window, sample covariance, data_range 255) on uint8 HWC pairs.
// HBM-bound byte work: one pass over the image, coalesced along (x, c); the integer window sums of SSIM are exact.
// Arithmetic that must reproduce torch bit for bit uses the _rn intrinsics (no FMA contraction).
#include "common.cuh"
#include "host_util.h"

namespace uegan {

struct IoGeom {
  void* data;
  int n, h, w, c, halo;
  long long wp, hp;
  int es;  // element size: 4 (fp32, tf32-rounded) or 2
  int dtype;
};

__device__ __forceinline__ float io_round_tf32(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

// one thread per padded pixel: 3 bytes in, one 16-byte channel vector out (+ three fp32 plane values for interior pixels)
__global__ void pack_input_u8_kernel(const uint8_t* __restrict__ src, IoGeom d, int reflect, float m0, float m1, float m2,
                                     float s0, float s1, float s2, float* __restrict__ planes, long long total) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int xp = (int)(i % d.wp);
  const int yp = (int)((i / d.wp) % d.hp);
  const int n = (int)(i / (d.wp * d.hp));
  int y = yp - d.halo, x = xp - d.halo;
  const bool in_halo = (y < 0) || (y >= d.h) || (x < 0) || (x >= d.w);
  float v0 = 0.f, v1 = 0.f, v2 = 0.f;
  if (!in_halo || reflect) {
    y = reflect_idx(y, d.h);
    x = reflect_idx(x, d.w);
    const uint8_t* s = src + (((long long)n * d.h + y) * d.w + x) * 3;
    // ToTensor: u8 -> float / 255 ; Normalize: (t - mean) / std   -- IEEE division, exactly torch's two steps
    v0 = __fdiv_rn(__fsub_rn(__fdiv_rn((float)s[0], 255.f), m0), s0);
    v1 = __fdiv_rn(__fsub_rn(__fdiv_rn((float)s[1], 255.f), m1), s1);
    v2 = __fdiv_rn(__fsub_rn(__fdiv_rn((float)s[2], 255.f), m2), s2);
    if (!in_halo && planes) {
      const long long plane = (long long)d.h * d.w;
      float* p = planes + (long long)n * 3 * plane + (long long)y * d.w + x;
      p[0] = v0; p[plane] = v1; p[2 * plane] = v2;
    }
  }
  if (d.data == nullptr) return;
  uint8_t* dp = static_cast<uint8_t*>(d.data) + i * d.c * d.es;
  if (d.es == 4) {
    *reinterpret_cast<float4*>(dp) = make_float4(io_round_tf32(v0), io_round_tf32(v1), io_round_tf32(v2), 0.f);
    for (int c = 4; c < d.c; c += 4) *reinterpret_cast<float4*>(dp + c * 4) = make_float4(0.f, 0.f, 0.f, 0.f);
  } else {
    uint32_t w0, w1;
    if (d.dtype == UEGAN_BF16) {
      __nv_bfloat162 a = __floats2bfloat162_rn(v0, v1), b = __floats2bfloat162_rn(v2, 0.f);
      w0 = *reinterpret_cast<uint32_t*>(&a); w1 = *reinterpret_cast<uint32_t*>(&b);
    } else {
      __half2 a = __floats2half2_rn(v0, v1), b = __floats2half2_rn(v2, 0.f);
      w0 = *reinterpret_cast<uint32_t*>(&a); w1 = *reinterpret_cast<uint32_t*>(&b);
    }
    *reinterpret_cast<uint4*>(dp) = make_uint4(w0, w1, 0u, 0u);
    for (int c = 8; c < d.c; c += 8) *reinterpret_cast<uint4*>(dp + c * 2) = make_uint4(0u, 0u, 0u, 0u);
  }
}

// fp32 NCHW in [-1, 1] -> uint8 HWC:  denorm = clamp((x + 1) / 2, 0, 1);  save_image: clamp(v * 255 + 0.5, 0, 255) -> uint8
__global__ void unpack_output_u8_kernel(const float* __restrict__ src, uint8_t* __restrict__ dst, int h, int w, int cch,
                                        long long total) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;  // output byte index (n, y, x, c)
  if (i >= total) return;
  const int c = (int)(i % cch);
  const long long pix = i / cch;
  const long long plane = (long long)h * w;
  const long long n = pix / plane, yx = pix % plane;
  float v = __ldg(src + (n * cch + c) * plane + yx);
  v = __fdiv_rn(__fadd_rn(v, 1.f), 2.f);
  v = fminf(fmaxf(v, 0.f), 1.f);
  v = __fadd_rn(__fmul_rn(v, 255.f), 0.5f);
  v = fminf(fmaxf(v, 0.f), 255.f);
  dst[i] = (uint8_t)v;  // truncation, as Tensor.to(torch.uint8)
}

// ------------------------------------------------------------------------------------------------------------------
// PSNR: exact integer sum of squared differences over the cropped region, per image
// ------------------------------------------------------------------------------------------------------------------
__global__ void sse_u8_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int h, int w, int cch, int crop,
                              unsigned long long* __restrict__ sse) {
  pdl_sync();
  const int n = blockIdx.y;
  const int rw = (w - 2 * crop) * cch;  // bytes per cropped row
  const int rows = h - 2 * crop;
  const long long total = (long long)rows * rw;
  unsigned long long acc = 0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int y = (int)(i / rw) + crop, xc = (int)(i % rw) + crop * cch;
    const long long o = ((long long)n * h + y) * w * cch + xc;
    const int dlt = (int)a[o] - (int)b[o];
    acc += (unsigned)(dlt * dlt);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0 && acc) atomicAdd(sse + n, acc);
}

// ------------------------------------------------------------------------------------------------------------------
// SSIM (skimage.metrics.structural_similarity, gaussian_weights=False, win_size=7, use_sample_covariance=True,
// data_range=255, multichannel): per channel
//   ux = mean_7x7(X) ...  vx = 49/48 (mean(XX) - ux^2), vxy = 49/48 (mean(XY) - ux uy)
//   S = (2 ux uy + C1)(2 vxy + C2) / ((ux^2 + uy^2 + C1)(vx + vy + C2)),  mssim = mean of S over pixels >= 3 from the border
// (the border crop makes the filter's boundary mode irrelevant).  Window sums are exact integers; the rest is fp64.
// One thread per (pixel, channel) of the valid region; fp64 block sums, one atomic per block.
// ------------------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) ssim_u8_kernel(const uint8_t* __restrict__ a, const uint8_t* __restrict__ b, int h, int w,
                                                     int cch, int crop, double* __restrict__ sums) {
  pdl_sync();
  __shared__ double sh[8];
  const int n = blockIdx.y;
  const int H = h - 2 * crop, W = w - 2 * crop;  // the metric runs on the cropped image (CalcSSIM.py:47-52)
  const int vh = H - 6, vw = W - 6;              // valid region after skimage's crop by (win_size - 1) / 2 = 3
  const long long total = (long long)vh * vw * cch;
  const double C1 = (0.01 * 255.0) * (0.01 * 255.0), C2 = (0.03 * 255.0) * (0.03 * 255.0);
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int c = (int)(i % cch);
    const int x = (int)((i / cch) % vw) + crop;  // top-left corner of the 7x7 window in full-image coordinates
    const int y = (int)(i / ((long long)cch * vw)) + crop;
    unsigned sx = 0, sy = 0, sxx = 0, syy = 0, sxy = 0;
    for (int r = 0; r < 7; ++r) {
      const long long o = (((long long)n * h + y + r) * w + x) * cch + c;
#pragma unroll
      for (int s = 0; s < 7; ++s) {
        const unsigned p = a[o + s * cch], q = b[o + s * cch];
        sx += p; sy += q; sxx += p * p; syy += q * q; sxy += p * q;
      }
    }
    const double ux = sx / 49.0, uy = sy / 49.0;
    const double cn = 49.0 / 48.0;
    const double vx = cn * (sxx / 49.0 - ux * ux), vy = cn * (syy / 49.0 - uy * uy), vxy = cn * (sxy / 49.0 - ux * uy);
    const double A1 = 2.0 * ux * uy + C1, A2 = 2.0 * vxy + C2, B1 = ux * ux + uy * uy + C1, B2 = vx + vy + C2;
    acc += (A1 * A2) / (B1 * B2);
  }
  for (int o = 16; o > 0; o >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, o);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double t = 0.0;
    for (int i = 0; i < 8; ++i) t += sh[i];
    atomicAdd(sums + n, t);
  }
}

}  // namespace uegan

using namespace uegan;

extern "C" {

int uegan_pack_input_u8(const uint8_t* img_nhwc_u8, int32_t n, int32_t h, int32_t w, const uegan_tensor* dst,
                        float* x_nchw_out, int32_t pad_mode, const float* mean_host, const float* std_host, void* stream) {
  UEGAN_CHECK(img_nhwc_u8 && mean_host && std_host && (dst || x_nchw_out), "pack_input_u8: null pointer");
  IoGeom d;
  memset(&d, 0, sizeof(d));
  d.n = n; d.h = h; d.w = w; d.wp = w; d.hp = h; d.c = 4; d.es = 4;
  if (dst) {
    UEGAN_CHECK(dst->data && dtype_ok(dst->dtype), "pack_input_u8: bad destination tensor");
    UEGAN_CHECK(dst->n == n && dst->h == h && dst->w == w, "pack_input_u8: destination is %dx%dx%d, image batch %dx%dx%d",
                dst->n, dst->h, dst->w, n, h, w);
    UEGAN_CHECK((dst->c * dtype_size(dst->dtype)) % 16 == 0 && dst->c >= 3, "pack_input_u8: bad channel count %d", dst->c);
    if (pad_mode == UEGAN_PAD_REFLECT)
      UEGAN_CHECK(dst->halo < h && dst->halo < w, "pack_input_u8: reflect halo %d needs h, w > halo", dst->halo);
    d.data = dst->data; d.c = dst->c; d.halo = dst->halo; d.wp = t_wp(*dst); d.hp = t_hp(*dst);
    d.es = dtype_size(dst->dtype); d.dtype = dst->dtype;
  }
  const long long total = (long long)n * d.hp * d.wp;
  if (total == 0) return 0;
  launch_pdl(pack_input_u8_kernel, (unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream), img_nhwc_u8, d, pad_mode == UEGAN_PAD_REFLECT, mean_host[0], mean_host[1], mean_host[2], std_host[0], std_host[1],
      std_host[2], x_nchw_out, total);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_unpack_output_u8(const float* x_nchw, uint8_t* out_nhwc_u8, int32_t n, int32_t c, int32_t h, int32_t w,
                           void* stream) {
  UEGAN_CHECK(x_nchw && out_nhwc_u8 && c >= 1, "unpack_output_u8: null pointer");
  const long long total = (long long)n * h * w * c;
  if (total == 0) return 0;
  launch_pdl(unpack_output_u8_kernel, (unsigned)((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream), x_nchw, out_nhwc_u8, h, w, c, total);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_sse_u8(const uint8_t* a_nhwc, const uint8_t* b_nhwc, int32_t n, int32_t h, int32_t w, int32_t c, int32_t crop,
                 uint64_t* sse_out, void* stream) {
  UEGAN_CHECK(a_nhwc && b_nhwc && sse_out, "sse_u8: null pointer");
  UEGAN_CHECK(crop >= 0 && h > 2 * crop && w > 2 * crop && c >= 1 && n >= 1, "sse_u8: empty cropped region");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  UEGAN_CUDA(cudaMemsetAsync(sse_out, 0, sizeof(uint64_t) * n, st));
  const long long total = (long long)(h - 2 * crop) * (w - 2 * crop) * c;
  long long blocks = (total + 256 * 8 - 1) / (256 * 8);
  if (blocks > 4 * num_sms()) blocks = 4 * num_sms();
  if (blocks < 1) blocks = 1;
  launch_pdl(sse_u8_kernel, dim3((unsigned)blocks, (unsigned)n), 256, 0, st, a_nhwc, b_nhwc, h, w, c, crop,
                                                                    reinterpret_cast<unsigned long long*>(sse_out));
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_ssim_u8(const uint8_t* a_nhwc, const uint8_t* b_nhwc, int32_t n, int32_t h, int32_t w, int32_t c, int32_t crop,
                  double* ssim_sum_out, void* stream) {
  UEGAN_CHECK(a_nhwc && b_nhwc && ssim_sum_out, "ssim_u8: null pointer");
  UEGAN_CHECK(crop >= 0 && h - 2 * crop >= 7 && w - 2 * crop >= 7 && c >= 1 && n >= 1,
              "ssim_u8: the cropped image must be at least 7x7 (win_size)");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  UEGAN_CUDA(cudaMemsetAsync(ssim_sum_out, 0, sizeof(double) * n, st));
  const long long total = (long long)(h - 2 * crop - 6) * (w - 2 * crop - 6) * c;
  long long blocks = (total + 255) / 256;
  if (blocks > 8 * num_sms()) blocks = 8 * num_sms();
  launch_pdl(ssim_u8_kernel, dim3((unsigned)blocks, (unsigned)n), 256, 0, st, a_nhwc, b_nhwc, h, w, c, crop, ssim_sum_out);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
