// HBM-bound helper kernels around the implicit-GEMM convolutions: layout packing, halo (padding) fill,
// InstanceNorm, bilinear x2 upsample, 2x2 max-pool.  All work on NHWC-with-halo tensors (uegan_sm100.h) in
// 16-byte channel vectors so that every global access is a full 16 B per thread, coalesced along C then W.
#include "common.cuh"
#include "host_util.h"

namespace uegan {

__device__ __forceinline__ float round_tf32_ew(float v) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(v));
  return __uint_as_float(r);
}

template <typename T> struct Vec;
template <> struct Vec<float> {
  static constexpr int N = 4;
  __device__ static void load(const float* p, float (&v)[4]) {
    const float4 q = *reinterpret_cast<const float4*>(p);
    v[0] = q.x; v[1] = q.y; v[2] = q.z; v[3] = q.w;
  }
  // F32 activations feed kind::tf32 MMAs: round to nearest tf32 once, at the producer.
  __device__ static void store(float* p, const float (&v)[4]) {
    *reinterpret_cast<float4*>(p) = make_float4(round_tf32_ew(v[0]), round_tf32_ew(v[1]), round_tf32_ew(v[2]),
                                                round_tf32_ew(v[3]));
  }
};
template <> struct Vec<__nv_bfloat16> {
  static constexpr int N = 8;
  __device__ static void load(const __nv_bfloat16* p, float (&v)[8]) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
      v[2 * i] = __low2float(h);
      v[2 * i + 1] = __high2float(h);
    }
  }
  __device__ static void store(__nv_bfloat16* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      __nv_bfloat162 h = __floats2bfloat162_rn(v[2 * i], v[2 * i + 1]);
      w[i] = *reinterpret_cast<uint32_t*>(&h);
    }
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};

template <> struct Vec<__half> {
  static constexpr int N = 8;
  __device__ static void load(const __half* p, float (&v)[8]) {
    const uint4 q = *reinterpret_cast<const uint4*>(p);
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
      v[2 * i] = __low2float(h);
      v[2 * i + 1] = __high2float(h);
    }
  }
  // saturating: a value beyond fp16's range is stored as +-65504, never as inf (a scale that lags by more than its 64x
  // margin then costs precision for one pass -- uegan_scale_update sees the saturated sample and backs off -- instead of
  // poisoning the step with inf - inf = nan)
  __device__ static void store(__half* p, const float (&v)[8]) {
    uint32_t w[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)  // F2FP.SATFINITE: one instruction per pair instead of four FMNMX + a convert
      asm("cvt.rn.satfinite.f16x2.f32 %0, %1, %2;" : "=r"(w[i]) : "f"(v[2 * i + 1]), "f"(v[2 * i]));
    *reinterpret_cast<uint4*>(p) = make_uint4(w[0], w[1], w[2], w[3]);
  }
};
// Raw 16-byte vector load + a separate convert: kernels with several optional operands issue EVERY load first and convert /
// combine afterwards, so one thread keeps all its operands in flight instead of waiting on each in turn (Vec<T>::load
// followed by its use inside `if (has_x)` serialised up to five global-memory latencies per thread, r4e).
__device__ __forceinline__ uint4 ldg16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }
template <typename T> __device__ __forceinline__ void cvt16(const uint4& q, float (&v)[Vec<T>::N]);
template <> __device__ __forceinline__ void cvt16<float>(const uint4& q, float (&v)[4]) {
  v[0] = __uint_as_float(q.x); v[1] = __uint_as_float(q.y); v[2] = __uint_as_float(q.z); v[3] = __uint_as_float(q.w);
}
template <> __device__ __forceinline__ void cvt16<__half>(const uint4& q, float (&v)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
    v[2 * i] = __low2float(h);
    v[2 * i + 1] = __high2float(h);
  }
}
template <> __device__ __forceinline__ void cvt16<__nv_bfloat16>(const uint4& q, float (&v)[8]) {
  const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
    v[2 * i] = __low2float(h);
    v[2 * i + 1] = __high2float(h);
  }
}
template <typename T> __device__ __forceinline__ float to_f32(T v);
template <> __device__ __forceinline__ float to_f32<float>(float v) { return v; }
template <> __device__ __forceinline__ float to_f32<__nv_bfloat16>(__nv_bfloat16 v) { return __bfloat162float(v); }
template <> __device__ __forceinline__ float to_f32<__half>(__half v) { return __half2float(v); }

struct TGeom {  // device-side view of a uegan_tensor
  void* data;
  int n, h, w, c, halo;
  long long wp, hp;
  const float* scale;  // uegan_tensor.scale: stored = scale * true (NULL = 1)
};
static TGeom geom(const uegan_tensor& t) {
  TGeom g;
  g.data = t.data; g.n = t.n; g.h = t.h; g.w = t.w; g.c = t.c; g.halo = t.halo;
  g.wp = t_wp(t); g.hp = t_hp(t);
  g.scale = t.scale;
  return g;
}
// the tensor's scale / its reciprocal (powers of two: exact)
__device__ __forceinline__ float tscale(const TGeom& g) { return g.scale ? __ldg(g.scale) : 1.f; }
// (the scale is a power of two by contract, include/uegan_sm100.h: its reciprocal is an exponent flip -- one integer
// subtraction instead of MUFU.RCP + the IEEE fix-up sequence that every THREAD of the elementwise kernels executed for up to
// five tensors: ~50 of grad_combine's ~200 instructions per 16-byte vector, r3r)
__device__ __forceinline__ float pow2_rcp(float s) { return __int_as_float(0x7F000000 - __float_as_int(s)); }
__device__ __forceinline__ float tinv(const TGeom& g) { return g.scale ? pow2_rcp(__ldg(g.scale)) : 1.f; }
// element offset of (n, y, x, c) with y, x in interior coordinates (may be negative into the halo)
// (row index in 32 bits -- n * hp + y never leaves them --, one widening multiply for the pixel index, one 64 x 32 multiply
// for the element offset: the all-64-bit form cost ~15 instructions per operand of every elementwise thread)
__device__ __forceinline__ long long toff(const TGeom& g, int n, int y, int x, int c) {
  const int row = n * (int)g.hp + y + g.halo;
  const long long pix = (long long)row * (int)g.wp + (x + g.halo);
  return pix * g.c + c;
}

// ------------------------------------------------------------------------------------------
// Strip reduction skeleton for every per-channel / per-(n, c) sum over pixels (InstanceNorm statistics and their
// backward, bias gradients, the perceptual MSE): block = (`rows` image rows of image n), thread = (pixel lane,
// 16-byte channel vector).  Loads are 16 B per thread, coalesced along C then W; integer division only per row;
// fp32 partials per thread, one smem tree per block, then Op::flush (atomics) once per (block, channel).
//   Op: void prep(int n, int c)                                     -- cache what is constant per (n, channel vector)
//       void acc(int n, int y, int x, int c, float (&a)[VN][NACC])  -- add the contributions of channels c..c+VN-1
//       void flush(int n, int c, const float (&tot)[NACC])          -- publish one channel's block total
// ------------------------------------------------------------------------------------------
template <typename T, int NACC, typename Op>
__global__ void __launch_bounds__(256) strip_reduce_kernel(Op op, int cch, int h, int w, int rows) {
  pdl_sync();
  constexpr int VN = Vec<T>::N;
  __shared__ float sh[256 * VN * NACC];
  const int cv = cch / VN;
  const int lanes = 256 / cv;
  const int lane = threadIdx.x / cv, c = (threadIdx.x % cv) * VN;
  const int n = blockIdx.y;
  const int y0 = blockIdx.x * rows, y1 = min(y0 + rows, h);
  float a[VN][NACC];
#pragma unroll
  for (int k = 0; k < VN; ++k)
#pragma unroll
    for (int j = 0; j < NACC; ++j) a[k][j] = 0.f;
  if (lane < lanes) {
    op.prep(n, c);  // per-(n, c) statistics of this thread's channel vector -> registers, once
    for (int y = y0; y < y1; ++y) {
      // four independent pixels per trip: the loads of a trip are issued back to back (the loop carries no stores), which
      // is what a latency-bound 16-byte-per-thread reduction needs (1.9-2.7 TB/s -> HBM-bound)
      int x = lane;
      // (16-bit tensors only: measured r1i, the fp32 reductions lose occupancy to the extra registers)
      for (; sizeof(T) == 2 && x + 3 * lanes < w; x += 4 * lanes) {
        op.acc(n, y, x, c, a);
        op.acc(n, y, x + lanes, c, a);
        op.acc(n, y, x + 2 * lanes, c, a);
        op.acc(n, y, x + 3 * lanes, c, a);
      }
      for (; x < w; x += lanes) op.acc(n, y, x, c, a);
    }
  }
#pragma unroll
  for (int k = 0; k < VN; ++k)
#pragma unroll
    for (int j = 0; j < NACC; ++j) sh[(threadIdx.x * VN + k) * NACC + j] = (lane < lanes) ? a[k][j] : 0.f;
  __syncthreads();
  // each thread totals one channel (or several when cch > 256) over the pixel lanes
  for (int ch = threadIdx.x; ch < cch; ch += 256) {
    const int v = ch / VN, k = ch % VN;
    float tot[NACC];
#pragma unroll
    for (int j = 0; j < NACC; ++j) tot[j] = 0.f;
    for (int l = 0; l < lanes; ++l)
#pragma unroll
      for (int j = 0; j < NACC; ++j) tot[j] += sh[((l * cv + v) * VN + k) * NACC + j];
    op.flush(n, ch, tot);
  }
}
template <typename T, int NACC, typename Op>
static void launch_strip_reduce(const Op& op, int cch, int n, int h, int w, cudaStream_t st) {
  const int rows = h >= 64 ? 8 : (h >= 16 ? 4 : h);
  dim3 grid((unsigned)((h + rows - 1) / rows), (unsigned)n);
  launch_pdl(strip_reduce_kernel<T, NACC, Op>, grid, 256, 0, st, op, cch, h, w, rows);
}

// ------------------------------------------------------------------------------------------
// NCHW fp32 (3 channels) -> NHWC with halo
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void pack_input_kernel(const float* __restrict__ src, TGeom d, int reflect, float s0, float s1, float s2,
                                  float b0, float b1, float b2, long long total) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int xp = (int)(i % d.wp);
  const int yp = (int)((i / d.wp) % d.hp);
  const int n = (int)(i / (d.wp * d.hp));
  int y = yp - d.halo, x = xp - d.halo;
  const bool in_halo = (y < 0) || (y >= d.h) || (x < 0) || (x >= d.w);
  float v[Vec<T>::N];
#pragma unroll
  for (int k = 0; k < Vec<T>::N; ++k) v[k] = 0.f;
  if (!in_halo || reflect) {
    y = reflect_idx(y, d.h);
    x = reflect_idx(x, d.w);
    const long long plane = (long long)d.h * d.w;
    const float* s = src + (long long)n * 3 * plane + (long long)y * d.w + x;
    const float so = tscale(d);
    v[0] = (__ldg(s) * s0 + b0) * so;
    v[1] = (__ldg(s + plane) * s1 + b1) * so;
    v[2] = (__ldg(s + 2 * plane) * s2 + b2) * so;
  }
  T* dp = static_cast<T*>(d.data) + i * d.c;
  Vec<T>::store(dp, v);
  // channels beyond the first vector (if c > Vec::N) are zero
  float z[Vec<T>::N];
#pragma unroll
  for (int k = 0; k < Vec<T>::N; ++k) z[k] = 0.f;
  for (int c = Vec<T>::N; c < d.c; c += Vec<T>::N) Vec<T>::store(dp + c, z);
}

// ------------------------------------------------------------------------------------------
// halo fill: top/bottom bands (halo rows x full padded width) then left/right bands (h rows x halo cols)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void halo_fill_kernel(TGeom t, int reflect, long long total) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cv = t.c / Vec<T>::N;
  const int c = (int)(i % cv) * Vec<T>::N;
  long long pix = i / cv;
  const long long per_img = 2LL * t.halo * t.wp + 2LL * t.halo * t.h;
  const int n = (int)(pix / per_img);
  pix %= per_img;
  int yp, xp;
  if (pix < 2LL * t.halo * t.wp) {
    const int band_row = (int)(pix / t.wp);
    xp = (int)(pix % t.wp);
    yp = band_row < t.halo ? band_row : (int)(t.hp - 2 * t.halo + band_row);
  } else {
    pix -= 2LL * t.halo * t.wp;
    const int row = (int)(pix / (2 * t.halo));
    const int col = (int)(pix % (2 * t.halo));
    yp = t.halo + row;
    xp = col < t.halo ? col : (int)(t.wp - 2 * t.halo + col);
  }
  T* base = static_cast<T*>(t.data);
  float v[Vec<T>::N];
  if (reflect) {
    const int y = reflect_idx(yp - t.halo, t.h), x = reflect_idx(xp - t.halo, t.w);
    Vec<T>::load(base + toff(t, n, y, x, c), v);
  } else {
#pragma unroll
    for (int k = 0; k < Vec<T>::N; ++k) v[k] = 0.f;
  }
  Vec<T>::store(base + (((long long)n * t.hp + yp) * t.wp + xp) * t.c + c, v);
}

// ------------------------------------------------------------------------------------------
// InstanceNorm: per-(n,c) sum / sum of squares in fp64 (block partials -> atomics), finalize, apply
// ------------------------------------------------------------------------------------------
template <typename T>
struct InStatsOp {
  TGeom s;
  double* stats;
  __device__ void prep(int, int) {}
  __device__ void acc(int n, int y, int x, int c, float (&a)[Vec<T>::N][2]) const {
    float v[Vec<T>::N];
    Vec<T>::load(static_cast<const T*>(s.data) + toff(s, n, y, x, c), v);
#pragma unroll
    for (int k = 0; k < Vec<T>::N; ++k) { a[k][0] += v[k]; a[k][1] += v[k] * v[k]; }
  }
  __device__ void flush(int n, int c, const float (&t)[2]) const {
    atomicAdd(&stats[((long long)n * s.c + c) * 2], (double)t[0]);
    atomicAdd(&stats[((long long)n * s.c + c) * 2 + 1], (double)t[1]);
  }
};
template <typename T>
static void run_in_stats(const TGeom& s, double* stats, cudaStream_t st) {
  InStatsOp<T> op{s, stats};
  launch_strip_reduce<T, 2>(op, s.c, s.n, s.h, s.w, st);
}

// (mean, rstd) of the STORED values: with stored = s * true, mean_st = s * mean and rstd_st = 1 / sqrt(s^2 var + s^2 eps)
// = rstd / s, so that (stored - mean_st) * rstd_st is the true normalised value.
__global__ void in_finalize_kernel(const double* __restrict__ stats, float* __restrict__ mr, int total, double inv_npix,
                                   float eps, const float* __restrict__ src_scale) {
  pdl_sync();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const double s = src_scale ? (double)__ldg(src_scale) : 1.0;
  const double mean = stats[2 * i] * inv_npix;
  double var = stats[2 * i + 1] * inv_npix - mean * mean;  // biased variance (InstanceNorm2d)
  if (var < 0.0) var = 0.0;
  mr[2 * i] = (float)mean;
  mr[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps * s * s));
}

// grid = (x-chunks of a row, rows, images): no per-thread divisions (cv is a power of two)
template <typename T>
__global__ void in_apply_kernel(TGeom s, TGeom d, int dst_c_off, const float* __restrict__ mr, int cv_log2) {
  pdl_sync();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int x = t >> cv_log2;
  if (x >= s.w) return;
  const int c = (t & ((1 << cv_log2) - 1)) * Vec<T>::N;
  const int y = blockIdx.y, n = blockIdx.z;
  float v[Vec<T>::N];
  Vec<T>::load(static_cast<const T*>(s.data) + toff(s, n, y, x, c), v);
  // the (mean, rstd) pairs of this thread's channels as 16-byte loads: 2 * N scalar loads per 16 bytes of data made the
  // kernel L1-bound (r3q: l1tex 90 %, 2.7 TB/s)
  const float4* m4 = reinterpret_cast<const float4*>(mr + ((long long)n * s.c + c) * 2);
  const float so = tscale(d);
#pragma unroll
  for (int k = 0; k < Vec<T>::N; k += 2) {
    const float4 q = __ldg(m4 + (k >> 1));  // mean_k, rstd_k, mean_{k+1}, rstd_{k+1}
    v[k] = (v[k] - q.x) * (q.y * so);
    v[k + 1] = (v[k + 1] - q.z) * (q.w * so);
  }
  Vec<T>::store(static_cast<T*>(d.data) + toff(d, n, y, x, dst_c_off + c), v);
}

// ------------------------------------------------------------------------------------------
// bilinear x2, align_corners=True
// ------------------------------------------------------------------------------------------
// grid = (x-chunks of an output row, output rows, images): no per-thread divisions (cv is a power of two)
template <typename T>
__global__ void upsample2x_kernel(TGeom s, TGeom d, int dst_c_off, float sy, float sx, int cv_log2) {
  pdl_sync();
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int xo = t >> cv_log2;
  if (xo >= d.w) return;
  const int c = (t & ((1 << cv_log2) - 1)) * Vec<T>::N;
  const int yo = blockIdx.y, n = blockIdx.z;
  const float fy = sy * yo, fx = sx * xo;
  const int y0 = (int)fy, x0 = (int)fx;
  const float ly = fy - y0, lx = fx - x0;
  // one address, two strides (the kernel was issue-bound: r3q sm throughput 75 %, 256 instructions per thread -- four
  // 64-bit offsets and a 9-operation lerp per channel); the four bilinear weights carry the re-scale factor
  const T* p00 = static_cast<const T*>(s.data) + toff(s, n, y0, x0, c);
  const long long dx = x0 < s.w - 1 ? s.c : 0, dy = y0 < s.h - 1 ? s.wp * s.c : 0;
  float a[Vec<T>::N], b[Vec<T>::N], e[Vec<T>::N], f[Vec<T>::N], o[Vec<T>::N];
  Vec<T>::load(p00, a);
  Vec<T>::load(p00 + dx, b);
  Vec<T>::load(p00 + dy, e);
  Vec<T>::load(p00 + dy + dx, f);
  const float rs = tscale(d) * tinv(s);  // re-scale from the source's to the destination's power-of-two scale
  const float w00 = (1.f - ly) * (1.f - lx) * rs, w01 = (1.f - ly) * lx * rs, w10 = ly * (1.f - lx) * rs, w11 = ly * lx * rs;
#pragma unroll
  for (int k = 0; k < Vec<T>::N; ++k) o[k] = fmaf(w11, f[k], fmaf(w10, e[k], fmaf(w01, b[k], w00 * a[k])));
  Vec<T>::store(static_cast<T*>(d.data) + toff(d, n, yo, xo, dst_c_off + c), o);
}

// ------------------------------------------------------------------------------------------
// Decoder concat in ONE pass: cat[.., 0:C) = bilinear x2 (align_corners) of u, cat[.., C:2C) = InstanceNorm(z) from the
// finalised (mean, rstd) pairs (models.py:55-67: torch.cat([upsample(y), GAM(skip)], 1)).  Written as whole pixels: the
// two separate passes each wrote HALF of every 128-byte line of cat and ran at 1.8 - 2.7 TB/s (r3h).
// grid = (x-chunks of an output row, output rows, images); a thread owns one 16-byte vector of the 2C-channel pixel.
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void cat_build_kernel(TGeom u, TGeom z, TGeom d, const float* __restrict__ mr, float sy, float sx, int cv_log2) {
  pdl_sync();
  constexpr int VN = Vec<T>::N;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int xo = t >> cv_log2;
  if (xo >= d.w) return;
  const int c = (t & ((1 << cv_log2) - 1)) * VN;  // channel of cat
  const int yo = blockIdx.y, n = blockIdx.z;
  float o[VN];
  if (c < u.c) {
    // (the arithmetic of upsample2x_kernel, operation for operation: the two paths agree bit for bit)
    const float fy = sy * yo, fx = sx * xo;
    const int y0 = (int)fy, x0 = (int)fx;
    const float ly = fy - y0, lx = fx - x0;
    const T* p00 = static_cast<const T*>(u.data) + toff(u, n, y0, x0, c);
    const long long dx = x0 < u.w - 1 ? u.c : 0, dy = y0 < u.h - 1 ? u.wp * u.c : 0;
    float a[VN], b[VN], e[VN], f[VN];
    Vec<T>::load(p00, a);
    Vec<T>::load(p00 + dx, b);
    Vec<T>::load(p00 + dy, e);
    Vec<T>::load(p00 + dy + dx, f);
    const float rs = tscale(d) * tinv(u);
    const float w00 = (1.f - ly) * (1.f - lx) * rs, w01 = (1.f - ly) * lx * rs, w10 = ly * (1.f - lx) * rs, w11 = ly * lx * rs;
#pragma unroll
    for (int k = 0; k < VN; ++k) o[k] = fmaf(w11, f[k], fmaf(w10, e[k], fmaf(w01, b[k], w00 * a[k])));
  } else {
    const int cz = c - u.c;
    Vec<T>::load(static_cast<const T*>(z.data) + toff(z, n, yo, xo, cz), o);
    const float4* m4 = reinterpret_cast<const float4*>(mr + ((long long)n * z.c + cz) * 2);
    const float so = tscale(d);
#pragma unroll
    for (int k = 0; k < VN; k += 2) {
      const float4 q = __ldg(m4 + (k >> 1));
      o[k] = (o[k] - q.x) * (q.y * so);
      o[k + 1] = (o[k + 1] - q.z) * (q.w * so);
    }
  }
  Vec<T>::store(static_cast<T*>(d.data) + toff(d, n, yo, xo, c), o);
}

template <typename T>
__global__ void maxpool2x2_kernel(TGeom s, TGeom d, long long total) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int cv = s.c / Vec<T>::N;
  const int c = (int)(i % cv) * Vec<T>::N;
  long long pix = i / cv;
  const int xo = (int)(pix % d.w);
  pix /= d.w;
  const int yo = (int)(pix % d.h);
  const int n = (int)(pix / d.h);
  const T* base = static_cast<const T*>(s.data);
  float a[Vec<T>::N], b[Vec<T>::N], e[Vec<T>::N], f[Vec<T>::N], o[Vec<T>::N];
  Vec<T>::load(base + toff(s, n, 2 * yo, 2 * xo, c), a);
  Vec<T>::load(base + toff(s, n, 2 * yo, 2 * xo + 1, c), b);
  Vec<T>::load(base + toff(s, n, 2 * yo + 1, 2 * xo, c), e);
  Vec<T>::load(base + toff(s, n, 2 * yo + 1, 2 * xo + 1, c), f);
#pragma unroll
  for (int k = 0; k < Vec<T>::N; ++k) o[k] = fmaxf(fmaxf(a[k], b[k]), fmaxf(e[k], f[k]));
  Vec<T>::store(static_cast<T*>(d.data) + toff(d, n, yo, xo, c), o);
}

template <typename T>
__global__ void unpack_nchw_kernel(TGeom s, int c_off, int c_count, float* __restrict__ dst, long long total) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % s.w);
  long long r = i / s.w;
  const int y = (int)(r % s.h);
  r /= s.h;
  const int c = (int)(r % c_count);
  const int n = (int)(r / c_count);
  const T* base = static_cast<const T*>(s.data);
  dst[i] = to_f32<T>(base[toff(s, n, y, x, c_off + c)]);
}

static inline unsigned nblocks(long long total, int threads) { return (unsigned)((total + threads - 1) / threads); }

static int check_vec(const uegan_tensor& t, const char* who) {
  UEGAN_CHECK(t.data != nullptr, "%s: null tensor", who);
  UEGAN_CHECK(dtype_ok(t.dtype), "%s: bad dtype", who);
  UEGAN_CHECK((t.c * dtype_size(t.dtype)) % 16 == 0, "%s: c*elem must be a multiple of 16 B", who);
  return 0;
}

}  // namespace uegan

using namespace uegan;

extern "C" {

int uegan_pack_input(const float* x_nchw, const uegan_tensor* dst, int32_t pad_mode, const float* scale_host,
                     const float* shift_host, void* stream) {
  UEGAN_CHECK(x_nchw && dst, "pack_input: null pointer");
  if (check_vec(*dst, "pack_input")) return -1;
  const TGeom d = geom(*dst);
  const long long total = (long long)d.n * d.hp * d.wp;
  const float s0 = scale_host ? scale_host[0] : 1.f, s1 = scale_host ? scale_host[1] : 1.f,
              s2 = scale_host ? scale_host[2] : 1.f;
  const float b0 = shift_host ? shift_host[0] : 0.f, b1 = shift_host ? shift_host[1] : 0.f,
              b2 = shift_host ? shift_host[2] : 0.f;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dst->dtype == UEGAN_F32)
    launch_pdl(pack_input_kernel<float>, nblocks(total, 256), 256, 0, st, x_nchw, d, pad_mode == UEGAN_PAD_REFLECT, s0, s1, s2,
                                                                  b0, b1, b2, total);
  else if (dst->dtype == UEGAN_BF16)
    launch_pdl(pack_input_kernel<__nv_bfloat16>, nblocks(total, 256), 256, 0, st, x_nchw, d, pad_mode == UEGAN_PAD_REFLECT, s0,
                                                                          s1, s2, b0, b1, b2, total);
  else
    launch_pdl(pack_input_kernel<__half>, nblocks(total, 256), 256, 0, st, x_nchw, d, pad_mode == UEGAN_PAD_REFLECT, s0,
                                                                          s1, s2, b0, b1, b2, total);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_halo_fill(const uegan_tensor* t, int32_t pad_mode, void* stream) {
  UEGAN_CHECK(t, "halo_fill: null pointer");
  if (check_vec(*t, "halo_fill")) return -1;
  if (t->halo == 0) return 0;
  if (pad_mode == UEGAN_PAD_REFLECT)
    UEGAN_CHECK(t->halo < t->h && t->halo < t->w, "halo_fill: reflect halo %d needs h,w > halo (got %dx%d)", t->halo,
                t->h, t->w);
  const TGeom g = geom(*t);
  const int vn = 16 / dtype_size(t->dtype);
  const long long total = (long long)g.n * (2LL * g.halo * g.wp + 2LL * g.halo * g.h) * (g.c / vn);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (t->dtype == UEGAN_F32)
    launch_pdl(halo_fill_kernel<float>, nblocks(total, 256), 256, 0, st, g, pad_mode == UEGAN_PAD_REFLECT, total);
  else if (t->dtype == UEGAN_BF16)
    launch_pdl(halo_fill_kernel<__nv_bfloat16>, nblocks(total, 256), 256, 0, st, g, pad_mode == UEGAN_PAD_REFLECT, total);
  else
    launch_pdl(halo_fill_kernel<__half>, nblocks(total, 256), 256, 0, st, g, pad_mode == UEGAN_PAD_REFLECT, total);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

static int instance_norm_impl(const uegan_tensor* src, const uegan_tensor* dst, int32_t dst_c_off, float eps,
                              double* stats_ws, void* stream, bool compute_stats) {
  UEGAN_CHECK(src && dst && stats_ws, "instance_norm: null pointer");
  if (check_vec(*src, "instance_norm src") || check_vec(*dst, "instance_norm dst")) return -1;
  UEGAN_CHECK(src->dtype == dst->dtype && src->n == dst->n && src->h == dst->h && src->w == dst->w,
              "instance_norm: src/dst mismatch");
  UEGAN_CHECK(dst_c_off >= 0 && dst_c_off + src->c <= dst->c && dst_c_off % 8 == 0, "instance_norm: bad slice");
  UEGAN_CHECK(src->c <= 256 * (16 / dtype_size(src->dtype)) && 256 % (src->c / (16 / dtype_size(src->dtype))) == 0,
              "instance_norm: unsupported channel count %d", src->c);
  const TGeom s = geom(*src), d = geom(*dst);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nc = s.n * s.c;
  const long long npix = (long long)s.h * s.w;
  if (compute_stats) {
    UEGAN_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * nc, st));
    if (src->dtype == UEGAN_F32) run_in_stats<float>(s, stats_ws, st);
    else if (src->dtype == UEGAN_BF16) run_in_stats<__nv_bfloat16>(s, stats_ws, st);
    else run_in_stats<__half>(s, stats_ws, st);
  }
  float* mr = reinterpret_cast<float*>(stats_ws + 2 * nc);  // finalised (mean, rstd) live after the raw sums
  launch_pdl(in_finalize_kernel, nblocks(nc, 128), 128, 0, st, stats_ws, mr, nc, 1.0 / (double)npix, eps, src->scale);
  const int vn = 16 / dtype_size(src->dtype);
  const int cv = s.c / vn;
  int lg = 0;
  while ((1 << lg) < cv) ++lg;
  UEGAN_CHECK((1 << lg) == cv, "instance_norm: channels / vector width must be a power of two (got %d)", cv);
  UEGAN_CHECK(s.h <= 65535 && s.n <= 65535, "instance_norm: tensor too large for the launch grid");
  const dim3 grid(nblocks((long long)s.w * cv, 256), (unsigned)s.h, (unsigned)s.n);
  if (src->dtype == UEGAN_F32)
    launch_pdl(in_apply_kernel<float>, grid, 256, 0, st, s, d, dst_c_off, mr, lg);
  else if (src->dtype == UEGAN_BF16)
    launch_pdl(in_apply_kernel<__nv_bfloat16>, grid, 256, 0, st, s, d, dst_c_off, mr, lg);
  else
    launch_pdl(in_apply_kernel<__half>, grid, 256, 0, st, s, d, dst_c_off, mr, lg);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_instance_norm_stats(const uegan_tensor* src, float eps, double* stats_ws, int32_t sums_ready,
                              float** mean_rstd_out, void* stream) {
  UEGAN_CHECK(src && stats_ws && mean_rstd_out, "instance_norm_stats: null pointer");
  if (check_vec(*src, "instance_norm_stats")) return -1;
  UEGAN_CHECK(src->c <= 256 * (16 / dtype_size(src->dtype)) && 256 % (src->c / (16 / dtype_size(src->dtype))) == 0,
              "instance_norm_stats: unsupported channel count %d", src->c);
  const TGeom s = geom(*src);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nc = s.n * s.c;
  const long long npix = (long long)s.h * s.w;
  if (!sums_ready) {
    UEGAN_CUDA(cudaMemsetAsync(stats_ws, 0, sizeof(double) * 2 * nc, st));
    if (src->dtype == UEGAN_F32) run_in_stats<float>(s, stats_ws, st);
    else if (src->dtype == UEGAN_BF16) run_in_stats<__nv_bfloat16>(s, stats_ws, st);
    else run_in_stats<__half>(s, stats_ws, st);
  }
  float* mr = reinterpret_cast<float*>(stats_ws + 2 * nc);
  launch_pdl(in_finalize_kernel, nblocks(nc, 128), 128, 0, st, stats_ws, mr, nc, 1.0 / (double)npix, eps, src->scale);
  *mean_rstd_out = mr;
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_instance_norm(const uegan_tensor* src, const uegan_tensor* dst, int32_t dst_c_off, float eps,
                        double* stats_ws, void* stream) {
  return instance_norm_impl(src, dst, dst_c_off, eps, stats_ws, stream, true);
}

int uegan_instance_norm_apply(const uegan_tensor* src, const uegan_tensor* dst, int32_t dst_c_off, float eps,
                              double* stats_ws, void* stream) {
  return instance_norm_impl(src, dst, dst_c_off, eps, stats_ws, stream, false);
}

int uegan_upsample2x(const uegan_tensor* src, const uegan_tensor* dst, int32_t dst_c_off, void* stream) {
  UEGAN_CHECK(src && dst, "upsample2x: null pointer");
  if (check_vec(*src, "upsample2x src") || check_vec(*dst, "upsample2x dst")) return -1;
  UEGAN_CHECK(src->dtype == dst->dtype && src->n == dst->n && dst->h == 2 * src->h && dst->w == 2 * src->w,
              "upsample2x: src/dst mismatch");
  UEGAN_CHECK(dst_c_off >= 0 && dst_c_off + src->c <= dst->c && dst_c_off % 8 == 0, "upsample2x: bad slice");
  const TGeom s = geom(*src), d = geom(*dst);
  const float sy = d.h > 1 ? (float)(s.h - 1) / (float)(d.h - 1) : 0.f;
  const float sx = d.w > 1 ? (float)(s.w - 1) / (float)(d.w - 1) : 0.f;
  const int vn = 16 / dtype_size(src->dtype);
  const int cv = s.c / vn;
  int lg = 0;
  while ((1 << lg) < cv) ++lg;
  UEGAN_CHECK((1 << lg) == cv, "upsample2x: channels / vector width must be a power of two (got %d)", cv);
  UEGAN_CHECK(d.h <= 65535 && d.n <= 65535, "upsample2x: tensor too large for the launch grid");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid(nblocks((long long)d.w * cv, 256), (unsigned)d.h, (unsigned)d.n);
  if (src->dtype == UEGAN_F32)
    launch_pdl(upsample2x_kernel<float>, grid, 256, 0, st, s, d, dst_c_off, sy, sx, lg);
  else if (src->dtype == UEGAN_BF16)
    launch_pdl(upsample2x_kernel<__nv_bfloat16>, grid, 256, 0, st, s, d, dst_c_off, sy, sx, lg);
  else
    launch_pdl(upsample2x_kernel<__half>, grid, 256, 0, st, s, d, dst_c_off, sy, sx, lg);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_cat_build(const uegan_tensor* u, const uegan_tensor* z, const float* mean_rstd, const uegan_tensor* dst,
                    void* stream) {
  UEGAN_CHECK(u && z && dst && mean_rstd, "cat_build: null pointer");
  if (check_vec(*u, "cat_build u") || check_vec(*z, "cat_build z") || check_vec(*dst, "cat_build dst")) return -1;
  UEGAN_CHECK(u->dtype == dst->dtype && z->dtype == dst->dtype && u->n == dst->n && z->n == dst->n && dst->h == 2 * u->h &&
                  dst->w == 2 * u->w && z->h == dst->h && z->w == dst->w && u->c == z->c && dst->c == 2 * u->c,
              "cat_build: u (n,h/2,w/2,C), z (n,h,w,C) and dst (n,h,w,2C) expected");
  const TGeom gu = geom(*u), gz = geom(*z), gd = geom(*dst);
  const float sy = gd.h > 1 ? (float)(gu.h - 1) / (float)(gd.h - 1) : 0.f;
  const float sx = gd.w > 1 ? (float)(gu.w - 1) / (float)(gd.w - 1) : 0.f;
  const int vn = 16 / dtype_size(dst->dtype);
  const int cv = gd.c / vn;
  int lg = 0;
  while ((1 << lg) < cv) ++lg;
  UEGAN_CHECK((1 << lg) == cv && gu.c % vn == 0, "cat_build: channels / vector width must be a power of two (got %d)", cv);
  UEGAN_CHECK(gd.h <= 65535 && gd.n <= 65535, "cat_build: tensor too large for the launch grid");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid(nblocks((long long)gd.w * cv, 256), (unsigned)gd.h, (unsigned)gd.n);
  if (dst->dtype == UEGAN_F32) launch_pdl(cat_build_kernel<float>, grid, 256, 0, st, gu, gz, gd, mean_rstd, sy, sx, lg);
  else if (dst->dtype == UEGAN_BF16) launch_pdl(cat_build_kernel<__nv_bfloat16>, grid, 256, 0, st, gu, gz, gd, mean_rstd, sy, sx, lg);
  else launch_pdl(cat_build_kernel<__half>, grid, 256, 0, st, gu, gz, gd, mean_rstd, sy, sx, lg);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_maxpool2x2(const uegan_tensor* src, const uegan_tensor* dst, void* stream) {
  UEGAN_CHECK(src && dst, "maxpool2x2: null pointer");
  if (check_vec(*src, "maxpool2x2 src") || check_vec(*dst, "maxpool2x2 dst")) return -1;
  UEGAN_CHECK(src->dtype == dst->dtype && src->n == dst->n && dst->h == src->h / 2 && dst->w == src->w / 2 &&
                  dst->c == src->c,
              "maxpool2x2: src/dst mismatch");
  const TGeom s = geom(*src), d = geom(*dst);
  const int vn = 16 / dtype_size(src->dtype);
  const long long total = (long long)d.n * d.h * d.w * (s.c / vn);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (src->dtype == UEGAN_F32)
    launch_pdl(maxpool2x2_kernel<float>, nblocks(total, 256), 256, 0, st, s, d, total);
  else if (src->dtype == UEGAN_BF16)
    launch_pdl(maxpool2x2_kernel<__nv_bfloat16>, nblocks(total, 256), 256, 0, st, s, d, total);
  else
    launch_pdl(maxpool2x2_kernel<__half>, nblocks(total, 256), 256, 0, st, s, d, total);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_unpack_nchw(const uegan_tensor* src, int32_t c_off, int32_t c_count, float* dst_nchw, void* stream) {
  UEGAN_CHECK(src && dst_nchw, "unpack_nchw: null pointer");
  UEGAN_CHECK(c_off >= 0 && c_off + c_count <= src->c, "unpack_nchw: bad channel range");
  const TGeom s = geom(*src);
  const long long total = (long long)s.n * c_count * s.h * s.w;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (src->dtype == UEGAN_F32)
    launch_pdl(unpack_nchw_kernel<float>, nblocks(total, 256), 256, 0, st, s, c_off, c_count, dst_nchw, total);
  else if (src->dtype == UEGAN_BF16)
    launch_pdl(unpack_nchw_kernel<__nv_bfloat16>, nblocks(total, 256), 256, 0, st, s, c_off, c_count, dst_nchw, total);
  else
    launch_pdl(unpack_nchw_kernel<__half>, nblocks(total, 256), 256, 0, st, s, c_off, c_count, dst_nchw, total);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
