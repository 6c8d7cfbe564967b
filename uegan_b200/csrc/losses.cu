// Loss reductions (warp-shuffle, fp64 accumulation) and their gradients.
//   relativistic average GAN loss, 5 scales     losses.py:348-377 (rahinge / rals), summed as in :393-409
//   InstanceNorm + MSE per VGG tap              losses.py:18, 30-34
//   multi-scale L1 / L2 / smooth-L1             losses.py:202-231 (AvgPool2d(2) pyramid, weights 1, 1/2, 1/4)
#include "common.cuh"
#include "host_util.h"

namespace uegan {

constexpr int kMaxScales = 8;
struct GanArgs {
  const float* real[kMaxScales];
  const float* fake[kMaxScales];
  float* d_real[kMaxScales];
  float* d_fake[kMaxScales];
  long long count[kMaxScales];
  int nscales;
  int mode;    // 0 rahinge, 1 rals
  int for_d;   // 1: discriminator side, 0: generator side
  double count_mul;  // data-parallel world size: means and normalisation are over the GLOBAL batch
};

__device__ __forceinline__ void block_atomic_add(double v, double* dst) {
  v = warp_sum(v);
  if ((threadIdx.x & 31) == 0 && v != 0.0) atomicAdd(dst, v);
}

// ws[2*i] = sum real_i, ws[2*i+1] = sum fake_i
__global__ void gan_sums_kernel(GanArgs a, double* __restrict__ ws) {
  pdl_sync();
  const int i = blockIdx.y;
  double sr = 0.0, sf = 0.0;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < a.count[i]; j += (long long)gridDim.x * blockDim.x) {
    sr += a.real[i][j];
    sf += a.fake[i][j];
  }
  block_atomic_add(sr, ws + 2 * i);
  block_atomic_add(sf, ws + 2 * i + 1);
}

// per scale: ws2[4*i+0] = sum term(real), [1] = sum term(fake), [2] = sum dterm/dx(real), [3] = sum dterm/dx(fake)
//  rahinge D: relu(1 - (r - mf)) , relu(1 + (f - mr));  G: relu(1 + (r - mf)), relu(1 - (f - mr))
//  rals    D: ((r - mf) - 1)^2   , ((f - mr) + 1)^2  ;  G: ((r - mf) + 1)^2 , ((f - mr) - 1)^2
__device__ __forceinline__ void gan_terms(int mode, float sgn, float x, float& term, float& dterm) {
  // term(x) with x = r - mean(f) (or f - mean(r)); sgn = +1 -> "push x below -1" form relu(1 + x) / (x + 1)^2,
  // sgn = -1 -> relu(1 - x) / (x - 1)^2
  const float z = 1.f + sgn * x;
  if (mode == 0) {
    term = fmaxf(z, 0.f);
    dterm = z > 0.f ? sgn : 0.f;
  } else {
    term = z * z;           // (x + sgn)^2 == (1 + sgn*x)^2
    dterm = 2.f * z * sgn;
  }
}
__global__ void gan_terms_kernel(GanArgs a, const double* __restrict__ ws, double* __restrict__ ws2) {
  pdl_sync();
  const int i = blockIdx.y;
  const double inv = 1.0 / ((double)a.count[i] * a.count_mul);
  const float mr = (float)(ws[2 * i] * inv), mf = (float)(ws[2 * i + 1] * inv);
  const float sr = a.for_d ? -1.f : 1.f;  // sign for the real-side term
  const float sf = -sr;
  double t0 = 0.0, t1 = 0.0, g0 = 0.0, g1 = 0.0;
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < a.count[i]; j += (long long)gridDim.x * blockDim.x) {
    float t, g;
    gan_terms(a.mode, sr, a.real[i][j] - mf, t, g);
    t0 += t; g0 += g;
    gan_terms(a.mode, sf, a.fake[i][j] - mr, t, g);
    t1 += t; g1 += g;
  }
  block_atomic_add(t0, ws2 + 4 * i);
  block_atomic_add(t1, ws2 + 4 * i + 1);
  block_atomic_add(g0, ws2 + 4 * i + 2);
  block_atomic_add(g1, ws2 + 4 * i + 3);
}
__global__ void gan_finalize_kernel(GanArgs a, const double* __restrict__ ws2, float* __restrict__ loss) {
  pdl_sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    double l = 0.0;
    for (int i = 0; i < a.nscales; ++i) l += 0.5 * (ws2[4 * i] + ws2[4 * i + 1]) / ((double)a.count[i] * a.count_mul);
    loss[0] = (float)l;
  }
}
// d loss / d real_j = gscale * 0.5/n * ( dterm_r(j) - mean_k dterm_f(k) ),  same for fake with roles swapped
// (the second part is the path through the batch-global mean, losses.py:351-353)
__global__ void gan_backward_kernel(GanArgs a, const double* __restrict__ ws, const double* __restrict__ ws2,
                                    const float* __restrict__ gscale_ptr, float gscale_host) {
  pdl_sync();
  const int i = blockIdx.y;
  const double inv = 1.0 / ((double)a.count[i] * a.count_mul);
  const float mr = (float)(ws[2 * i] * inv), mf = (float)(ws[2 * i + 1] * inv);
  const float sr = a.for_d ? -1.f : 1.f, sf = -sr;
  const float gs = (gscale_ptr ? gscale_ptr[0] : 1.f) * gscale_host * 0.5f * (float)inv;
  const float mean_gr = (float)(ws2[4 * i + 2] * inv), mean_gf = (float)(ws2[4 * i + 3] * inv);
  for (long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x; j < a.count[i]; j += (long long)gridDim.x * blockDim.x) {
    float t, g;
    if (a.d_real[i]) {
      gan_terms(a.mode, sr, a.real[i][j] - mf, t, g);
      a.d_real[i][j] += gs * (g - mean_gf);
    }
    if (a.d_fake[i]) {
      gan_terms(a.mode, sf, a.fake[i][j] - mr, t, g);
      a.d_fake[i][j] += gs * (g - mean_gr);
    }
  }
}

// ------------------------------------------------------------------------------------------
// InstanceNorm + MSE of one VGG tap.  mrx / mry: per-(n,c) (mean, rstd) float pairs.
// ------------------------------------------------------------------------------------------
struct TapGeom {
  const void* x;
  const void* y;
  int n, h, w, c, halo;
  long long wp, hp;
  int dtype;
};
__device__ __forceinline__ long long tap_off(const TapGeom& g, int n, int yy, int xx, int c) {
  return (((long long)n * g.hp + (yy + g.halo)) * g.wp + (xx + g.halo)) * g.c + c;
}
__device__ __forceinline__ float tap_load(const void* p, long long o, int dtype) {
  if (dtype == UEGAN_BF16) return __bfloat162float(static_cast<const __nv_bfloat16*>(p)[o]);
  if (dtype == UEGAN_F16) return __half2float(static_cast<const __half*>(p)[o]);
  return static_cast<const float*>(p)[o];
}
// accum[0] += sum ((x-mx)*rx - (y-my)*ry)^2   (strip-reduce skeleton of elementwise.cu: 16-byte loads, one atomic per
// (block, channel))
template <typename T>
struct InMseOp {
  TGeom x, y;
  const float* mrx;
  const float* mry;
  double* accum;
  float mx[Vec<T>::N], rx[Vec<T>::N], my[Vec<T>::N], ry[Vec<T>::N];
  __device__ void prep(int n, int c) {
#pragma unroll
    for (int k = 0; k < Vec<T>::N; ++k) {
      const long long si = ((long long)n * x.c + c + k) * 2;
      mx[k] = mrx[si]; rx[k] = mrx[si + 1]; my[k] = mry[si]; ry[k] = mry[si + 1];
    }
  }
  __device__ void acc(int n, int yy, int xx, int c, float (&a)[Vec<T>::N][1]) const {
    float xv[Vec<T>::N], yv[Vec<T>::N];
    Vec<T>::load(static_cast<const T*>(x.data) + toff(x, n, yy, xx, c), xv);
    Vec<T>::load(static_cast<const T*>(y.data) + toff(y, n, yy, xx, c), yv);
#pragma unroll
    for (int k = 0; k < Vec<T>::N; ++k) {
      const float d = (xv[k] - mx[k]) * rx[k] - (yv[k] - my[k]) * ry[k];
      a[k][0] += d * d;
    }
  }
  __device__ void flush(int n, int c, const float (&t)[1]) const { atomicAdd(accum, (double)t[0]); }
};
// ------------------------------------------------------------------------------------------
// VGG tap, joint pass: the five raw moments per (n, c) of the two feature maps -- sum x, sum x^2, sum y, sum y^2, sum xy --
// in ONE read of x and y.  Everything the tap needs follows in closed form (in_joint_finalize_kernel): the InstanceNorm
// statistics of both maps, the MSE of the normalised maps, and the two per-(n, c) sums of the backward pass.  Replaces
// four passes (statistics of x, of y, the MSE pass, the backward statistics pass: 6 tensor reads) by one (2 reads).
// ------------------------------------------------------------------------------------------
template <typename T>
struct JointMomentsOp {
  TGeom x, y;
  double* mom;
  __device__ void prep(int, int) {}
  __device__ void acc(int n, int yy, int xx, int c, float (&a)[Vec<T>::N][5]) const {
    float xv[Vec<T>::N], yv[Vec<T>::N];
    Vec<T>::load(static_cast<const T*>(x.data) + toff(x, n, yy, xx, c), xv);
    Vec<T>::load(static_cast<const T*>(y.data) + toff(y, n, yy, xx, c), yv);
#pragma unroll
    for (int k = 0; k < Vec<T>::N; ++k) {
      a[k][0] += xv[k];
      a[k][1] = fmaf(xv[k], xv[k], a[k][1]);
      a[k][2] += yv[k];
      a[k][3] = fmaf(yv[k], yv[k], a[k][3]);
      a[k][4] = fmaf(xv[k], yv[k], a[k][4]);
    }
  }
  __device__ void flush(int n, int c, const float (&t)[5]) const {
    double* m = mom + ((long long)n * x.c + c) * 5;
#pragma unroll
    for (int j = 0; j < 5; ++j) atomicAdd(m + j, (double)t[j]);
  }
};
// One thread per (n, c).  With N pixels and the fp32-rounded (mean, rstd) pairs the apply pass will use (mx, rx, my, ry):
//   sum xh     = rx (Sx - N mx)                      sum xh^2  = rx^2 (Sxx - 2 mx Sx + N mx^2)
//   sum xh yh  = rx ry (Sxy - mx Sy - my Sx + N mx my)
//   sums[0] = sum (xh - yh), sums[1] = sum (xh - yh) xh      (what TapBwdStatsOp measured in its own pass)
//   accum  += sum (xh - yh)^2 = sum xh^2 + sum yh^2 - 2 sum xh yh
__global__ void __launch_bounds__(256) in_joint_finalize_kernel(const double* __restrict__ mom, int nc, double npix, float eps,
                                                                const float* __restrict__ src_scale, float* __restrict__ mrx,
                                                                float* __restrict__ mry, double* __restrict__ sums,
                                                                double* __restrict__ accum) {
  pdl_sync();
  __shared__ double sh[256];
  const int i = blockIdx.x * 256 + threadIdx.x;
  double contrib = 0.0;
  if (i < nc) {
    const double s = src_scale ? (double)__ldg(src_scale) : 1.0;
    const double e = (double)eps * s * s;
    const double Sx = mom[5 * i], Sxx = mom[5 * i + 1], Sy = mom[5 * i + 2], Syy = mom[5 * i + 3], Sxy = mom[5 * i + 4];
    const double inv = 1.0 / npix;
    double vx = Sxx * inv - (Sx * inv) * (Sx * inv), vy = Syy * inv - (Sy * inv) * (Sy * inv);
    if (vx < 0.0) vx = 0.0;
    if (vy < 0.0) vy = 0.0;
    const float mxf = (float)(Sx * inv), rxf = (float)(1.0 / sqrt(vx + e));
    const float myf = (float)(Sy * inv), ryf = (float)(1.0 / sqrt(vy + e));
    mrx[2 * i] = mxf; mrx[2 * i + 1] = rxf;
    mry[2 * i] = myf; mry[2 * i + 1] = ryf;
    const double mx = mxf, rx = rxf, my = myf, ry = ryf;
    const double sxh = rx * (Sx - npix * mx), syh = ry * (Sy - npix * my);
    const double sxx = rx * rx * (Sxx - 2.0 * mx * Sx + npix * mx * mx);
    const double syy = ry * ry * (Syy - 2.0 * my * Sy + npix * my * my);
    const double sxy = rx * ry * (Sxy - mx * Sy - my * Sx + npix * mx * my);
    sums[2 * i] = sxh - syh;
    sums[2 * i + 1] = sxx - sxy;
    contrib = sxx + syy - 2.0 * sxy;
  }
  sh[threadIdx.x] = contrib;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) sh[threadIdx.x] += sh[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) atomicAdd(accum, sh[0]);
}
// loss += weight * accum / numel ; accum reset
__global__ void scalar_axpy_kernel(double* accum, double scale, float* loss) {
  pdl_sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    loss[0] += (float)(accum[0] * scale);
    accum[0] = 0.0;
  }
}

// ------------------------------------------------------------------------------------------
// multi-scale reconstruction loss on fp32 NCHW images; one thread per 4x4 pixel block of one (n, c) plane
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float rec_term(float d, int type) {
  if (type == 0) return fabsf(d);
  if (type == 2) return d * d;
  const float a = fabsf(d);  // smooth-L1, beta = 1
  return a < 1.f ? 0.5f * d * d : a - 0.5f;
}
__device__ __forceinline__ float rec_dterm(float d, int type) {
  if (type == 0) return d > 0.f ? 1.f : (d < 0.f ? -1.f : 0.f);
  if (type == 2) return 2.f * d;
  return fabsf(d) < 1.f ? d : (d > 0.f ? 1.f : -1.f);
}
// accum[0..2] = sum of term at scale 0, 1, 2.  If grad != NULL also writes d loss / d pred (needs w0..w2 = weight_i / N_i
// already multiplied by the upstream scale).
__global__ void msrec_kernel(const float* __restrict__ pred, const float* __restrict__ gt, int planes, int h, int w,
                             int type, int scales, double* __restrict__ accum, float* __restrict__ grad, float w0_,
                             float w1_, float w2_, const float* __restrict__ gscale) {
  pdl_sync();
  const float gs = gscale ? gscale[0] : 1.f;
  const float w0 = w0_ * gs, w1 = w1_ * gs, w2 = w2_ * gs;
  const int bw = w >> 2, bh = h >> 2;
  const long long total = (long long)planes * bh * bw;
  double s0 = 0.0, s1 = 0.0, s2 = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
    const int bx = (int)(i % bw);
    const int by = (int)((i / bw) % bh);
    const long long pl = i / ((long long)bw * bh);
    const long long base = pl * h * w + (long long)(by * 4) * w + bx * 4;
    float d[4][4];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const float4 p4 = *reinterpret_cast<const float4*>(pred + base + (long long)r * w);
      const float4 g4 = *reinterpret_cast<const float4*>(gt + base + (long long)r * w);
      d[r][0] = p4.x - g4.x; d[r][1] = p4.y - g4.y; d[r][2] = p4.z - g4.z; d[r][3] = p4.w - g4.w;
    }
    float d1[2][2], d2 = 0.f;
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        d1[r][c] = 0.25f * (d[2 * r][2 * c] + d[2 * r][2 * c + 1] + d[2 * r + 1][2 * c] + d[2 * r + 1][2 * c + 1]);
        d2 += 0.25f * d1[r][c];
      }
    float t0 = 0.f, t1 = 0.f;
#pragma unroll
    for (int r = 0; r < 4; ++r)
#pragma unroll
      for (int c = 0; c < 4; ++c) t0 += rec_term(d[r][c], type);
#pragma unroll
    for (int r = 0; r < 2; ++r)
#pragma unroll
      for (int c = 0; c < 2; ++c) t1 += rec_term(d1[r][c], type);
    s0 += t0;
    if (scales > 1) s1 += t1;
    if (scales > 2) s2 += rec_term(d2, type);
    if (grad) {
      const float g2 = scales > 2 ? w2 * rec_dterm(d2, type) * (1.f / 16.f) : 0.f;
#pragma unroll
      for (int r = 0; r < 4; ++r) {
        float o[4];
#pragma unroll
        for (int c = 0; c < 4; ++c) {
          float g = w0 * rec_dterm(d[r][c], type) + g2;
          if (scales > 1) g += w1 * rec_dterm(d1[r >> 1][c >> 1], type) * 0.25f;
          o[c] = g;
        }
        *reinterpret_cast<float4*>(grad + base + (long long)r * w) = make_float4(o[0], o[1], o[2], o[3]);
      }
    }
  }
  block_atomic_add(s0, accum);
  block_atomic_add(s1, accum + 1);
  block_atomic_add(s2, accum + 2);
}
__global__ void msrec_finalize_kernel(double* accum, double c0, double c1, double c2, float* loss) {
  pdl_sync();
  if (threadIdx.x == 0 && blockIdx.x == 0) {
    loss[0] = (float)(accum[0] * c0 + accum[1] * c1 + accum[2] * c2);
  }
}

}  // namespace uegan

using namespace uegan;

extern "C" {

static int fill_gan_args(GanArgs& a, int nscales, const float* const* real, const float* const* fake,
                         const int64_t* counts, int mode, int for_d) {
  UEGAN_CHECK(nscales >= 1 && nscales <= kMaxScales, "gan loss: bad number of scales %d", nscales);
  UEGAN_CHECK(mode == 0 || mode == 1, "gan loss: mode must be 0 (rahinge) or 1 (rals)");
  memset(&a, 0, sizeof(a));
  a.nscales = nscales; a.mode = mode; a.for_d = for_d; a.count_mul = 1.0;
  for (int i = 0; i < nscales; ++i) {
    UEGAN_CHECK(real[i] && fake[i] && counts[i] > 0, "gan loss: null / empty prediction map %d", i);
    a.real[i] = real[i]; a.fake[i] = fake[i]; a.count[i] = counts[i];
  }
  return 0;
}

int uegan_gan_loss_fwd(int32_t mode, int32_t for_discriminator, int32_t nscales, const float* const* real,
                       const float* const* fake, const int64_t* counts, double* ws, float* loss_out, void* stream) {
  GanArgs a;
  if (fill_gan_args(a, nscales, real, fake, counts, mode, for_discriminator)) return -1;
  UEGAN_CHECK(ws && loss_out, "gan loss: null workspace");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  UEGAN_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 6 * kMaxScales, st));
  dim3 grid(64, nscales);
  launch_pdl(gan_sums_kernel, grid, 256, 0, st, a, ws);
  launch_pdl(gan_terms_kernel, grid, 256, 0, st, a, ws, ws + 2 * kMaxScales);
  launch_pdl(gan_finalize_kernel, 1, 32, 0, st, a, ws + 2 * kMaxScales, loss_out);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_gan_loss_phase(int32_t phase, int32_t mode, int32_t for_discriminator, int32_t nscales,
                         const float* const* real, const float* const* fake, const int64_t* counts, int32_t world,
                         double* ws, float* loss_out, void* stream) {
  GanArgs a;
  if (fill_gan_args(a, nscales, real, fake, counts, mode, for_discriminator)) return -1;
  UEGAN_CHECK(ws && world >= 1 && phase >= 0 && phase <= 2, "gan loss phase: bad arguments");
  a.count_mul = (double)world;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  dim3 grid(64, nscales);
  if (phase == 0) {
    UEGAN_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 6 * kMaxScales, st));
    launch_pdl(gan_sums_kernel, grid, 256, 0, st, a, ws);
  } else if (phase == 1) {
    launch_pdl(gan_terms_kernel, grid, 256, 0, st, a, ws, ws + 2 * kMaxScales);
  } else {
    UEGAN_CHECK(loss_out, "gan loss phase 2: null loss");
    launch_pdl(gan_finalize_kernel, 1, 32, 0, st, a, ws + 2 * kMaxScales, loss_out);
  }
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_gan_loss_bwd(int32_t mode, int32_t for_discriminator, int32_t nscales, const float* const* real,
                       const float* const* fake, const int64_t* counts, const double* ws, float* const* d_real,
                       float* const* d_fake, const float* gscale_dev, float gscale_host, void* stream) {
  GanArgs a;
  if (fill_gan_args(a, nscales, real, fake, counts, mode, for_discriminator)) return -1;
  a.count_mul = gscale_host < 0.f ? (double)(-gscale_host) : 1.0;  // world size is passed as a negative gscale_host
  if (gscale_host < 0.f) gscale_host = 1.f;
  for (int i = 0; i < nscales; ++i) {
    a.d_real[i] = d_real ? d_real[i] : nullptr;
    a.d_fake[i] = d_fake ? d_fake[i] : nullptr;
  }
  dim3 grid(64, nscales);
  launch_pdl(gan_backward_kernel, grid, 256, 0, static_cast<cudaStream_t>(stream), a, ws, ws + 2 * kMaxScales, gscale_dev,
                                                                          gscale_host);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_in_mse_fwd(const uegan_tensor* x, const uegan_tensor* y, const float* mean_rstd_x, const float* mean_rstd_y,
                     float weight, double* accum, float* loss_inout, void* stream) {
  UEGAN_CHECK(x && y && mean_rstd_x && mean_rstd_y && accum && loss_inout, "in_mse: null pointer");
  UEGAN_CHECK(x->n == y->n && x->h == y->h && x->w == y->w && x->c == y->c && x->halo == y->halo &&
                  x->dtype == y->dtype,
              "in_mse: x / y mismatch");
  const TGeom gx = geom(*x), gy = geom(*y);
  const long long total = (long long)gx.n * gx.h * gx.w * gx.c;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int vn = 16 / dtype_size(x->dtype);
  UEGAN_CHECK(gx.c % vn == 0 && 256 % (gx.c / vn) == 0, "in_mse: unsupported channel count %d", gx.c);
  if (x->dtype == UEGAN_F32) {
    InMseOp<float> op{gx, gy, mean_rstd_x, mean_rstd_y, accum};
    launch_strip_reduce<float, 1>(op, gx.c, gx.n, gx.h, gx.w, st);
  } else if (x->dtype == UEGAN_BF16) {
    InMseOp<__nv_bfloat16> op{gx, gy, mean_rstd_x, mean_rstd_y, accum};
    launch_strip_reduce<__nv_bfloat16, 1>(op, gx.c, gx.n, gx.h, gx.w, st);
  } else {
    InMseOp<__half> op{gx, gy, mean_rstd_x, mean_rstd_y, accum};
    launch_strip_reduce<__half, 1>(op, gx.c, gx.n, gx.h, gx.w, st);
  }
  launch_pdl(scalar_axpy_kernel, 1, 32, 0, st, accum, (double)weight / (double)total, loss_inout);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_in_mse_joint(const uegan_tensor* x, const uegan_tensor* y, float eps, float weight, double* ws, double* accum,
                       float* loss_inout, float** mrx_out, float** mry_out, double** sums_out, void* stream) {
  UEGAN_CHECK(x && y && ws && accum && loss_inout && mrx_out && mry_out && sums_out, "in_mse_joint: null pointer");
  UEGAN_CHECK(x->n == y->n && x->h == y->h && x->w == y->w && x->c == y->c && x->dtype == y->dtype && x->scale == y->scale,
              "in_mse_joint: x / y mismatch");
  const TGeom gx = geom(*x), gy = geom(*y);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int vn = 16 / dtype_size(x->dtype);
  UEGAN_CHECK(gx.c % vn == 0 && 256 % (gx.c / vn) == 0, "in_mse_joint: unsupported channel count %d", gx.c);
  const int nc = gx.n * gx.c;
  const double npix = (double)gx.h * gx.w;
  double* mom = ws;
  double* sums = ws + 5 * (size_t)nc;
  float* mrx = reinterpret_cast<float*>(ws + 7 * (size_t)nc);
  float* mry = mrx + 2 * (size_t)nc;
  UEGAN_CUDA(cudaMemsetAsync(mom, 0, sizeof(double) * 5 * nc, st));
  if (x->dtype == UEGAN_F32) {
    JointMomentsOp<float> op{gx, gy, mom};
    launch_strip_reduce<float, 5>(op, gx.c, gx.n, gx.h, gx.w, st);
  } else if (x->dtype == UEGAN_BF16) {
    JointMomentsOp<__nv_bfloat16> op{gx, gy, mom};
    launch_strip_reduce<__nv_bfloat16, 5>(op, gx.c, gx.n, gx.h, gx.w, st);
  } else {
    JointMomentsOp<__half> op{gx, gy, mom};
    launch_strip_reduce<__half, 5>(op, gx.c, gx.n, gx.h, gx.w, st);
  }
  launch_pdl(in_joint_finalize_kernel, (nc + 255) / 256, 256, 0, st, mom, nc, npix, eps, x->scale, mrx, mry, sums, accum);
  launch_pdl(scalar_axpy_kernel, 1, 32, 0, st, accum, (double)weight / ((double)nc * npix), loss_inout);
  UEGAN_CUDA(cudaGetLastError());
  *mrx_out = mrx; *mry_out = mry; *sums_out = sums;
  return 0;
}

int uegan_msrec_loss(const float* pred_nchw, const float* gt_nchw, int32_t n, int32_t c, int32_t h, int32_t w,
                     int32_t type, int32_t scales, double* accum, float* loss_out, float* grad_nchw,
                     float grad_scale, const float* gscale_dev, void* stream) {
  UEGAN_CHECK(pred_nchw && gt_nchw && accum && loss_out, "msrec: null pointer");
  UEGAN_CHECK(h % 4 == 0 && w % 4 == 0, "msrec: H, W must be multiples of 4 (got %dx%d)", h, w);
  UEGAN_CHECK(scales >= 1 && scales <= 3 && type >= 0 && type <= 2, "msrec: bad scales/type");
  const double n0 = (double)n * c * h * w, n1 = n0 / 4, n2 = n0 / 16;
  const double c0 = 1.0 / n0, c1 = scales > 1 ? 0.5 / n1 : 0.0, c2 = scales > 2 ? 0.25 / n2 : 0.0;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  UEGAN_CUDA(cudaMemsetAsync(accum, 0, sizeof(double) * 3, st));
  const long long total = (long long)n * c * (h / 4) * (w / 4);
  int blocks = (int)((total + 255) / 256);
  if (blocks > 148 * 8) blocks = 148 * 8;
  launch_pdl(msrec_kernel, blocks, 256, 0, st, pred_nchw, gt_nchw, n * c, h, w, type, scales, accum, grad_nchw,
                                      (float)(c0 * grad_scale), (float)(c1 * grad_scale), (float)(c2 * grad_scale), gscale_dev);
  launch_pdl(msrec_finalize_kernel, 1, 32, 0, st, accum, c0, c1, c2, loss_out);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
