// HBM-bound kernels of the backward pass (everything that is not a GEMM): gradient of the planar heads, reflect-pad
// fold + activation masks, InstanceNorm backward, bilinear-x2 backward, max-pool backward, per-channel sums (bias
// gradients), InstanceNorm+MSE backward.  Same NHWC-with-halo tensors and 16-byte channel vectors as elementwise.cu.
// Gradient tensors that feed a dgrad GEMM are written WITH A ZERO HALO (the transposed convolution's padding).
#include "common.cuh"
#include "host_util.h"

namespace uegan {

// reuse Vec<T>, TGeom, geom(), toff(), to_f32<T>() from elementwise.cu (same translation unit)

// ------------------------------------------------------------------------------------------
// planar heads: dz[n,y,x,c] = dout_nchw * act'(.) (* clamp mask), written as NHWC with a zero halo
//   mode 0: out = tanh(z)                      -> dz = dout * (1 - out^2)                  (D heads, models.py:178)
//   mode 1: out = sigmoid(z)                   -> dz = dout * out * (1 - out)              (ls / rals heads)
//   mode 2: out = clamp(tanh(z) + x, -1, 1)    -> dz = dout * [|res + x| <= 1] * (1 - res^2), res given (models.py:72)
// ------------------------------------------------------------------------------------------
template <typename T>
__global__ void head_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ outv, const float* __restrict__ xin,
                                TGeom d, int cch, int mode, long long total) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int xp = (int)(i % d.wp);
  const int yp = (int)((i / d.wp) % d.hp);
  const int n = (int)(i / (d.wp * d.hp));
  const int y = yp - d.halo, x = xp - d.halo;
  float v[Vec<T>::N];
#pragma unroll
  for (int k = 0; k < Vec<T>::N; ++k) v[k] = 0.f;
  if (y >= 0 && y < d.h && x >= 0 && x < d.w) {
    const long long plane = (long long)d.h * d.w;
    const float so = tscale(d);
    for (int c = 0; c < cch; ++c) {
      const long long o = ((long long)n * cch + c) * plane + (long long)y * d.w + x;
      const float g = __ldg(dout + o), r = __ldg(outv + o);
      float dz;
      if (mode == 0) dz = g * (1.f - r * r);
      else if (mode == 1) dz = g * r * (1.f - r);
      else {
        const float s = r + __ldg(xin + o);
        dz = (s >= -1.f && s <= 1.f) ? g * (1.f - r * r) : 0.f;
      }
      v[c] = dz * so;
    }
  }
  T* dp = static_cast<T*>(d.data) + i * d.c;
  Vec<T>::store(dp, v);
  float z[Vec<T>::N];
#pragma unroll
  for (int k = 0; k < Vec<T>::N; ++k) z[k] = 0.f;
  for (int c = Vec<T>::N; c < d.c; c += Vec<T>::N) Vec<T>::store(dp + c, z);
}

// ------------------------------------------------------------------------------------------
// grad_combine: dst = mask(act; fwd) * mul * ( fold(src_a) + add_b + add_c ), interior; halo of dst := 0
//   src_a has spatial extent (h + 2*pa, w + 2*pa) (halo 0): gradient w.r.t. the PADDED input of a conv; reflect padding
//   folds the border back (adjoint of nn.ReflectionPad2d), zero padding crops.
// ------------------------------------------------------------------------------------------
struct CombineArgs {
  TGeom dst; int dst_c_off;
  TGeom a; int a_c_off; int pa; int reflect_a; int has_a;
  TGeom b; int b_c_off; int has_b;
  TGeom c; int c_c_off; int has_c;
  TGeom mask; int mask_c_off; int act; int has_mask;
  TGeom mul; int mul_c_off; int has_mul;
  // optional second output (halo 0): dst2 = mul2 * ( fold(src_a) + add_b + add_c ) -- the same sum before mask / mul
  // (the two factors of y4.mul(x1), models.py:70, get their gradients from one read of the incoming gradient)
  TGeom dst2; int dst2_c_off; int has_dst2;
  TGeom mul2; int mul2_c_off;
  int cch;  // channels produced
};
// grid = (x-chunks of a padded row, padded rows, images): no per-thread divisions (cv is a power of two)
template <typename T>
__global__ void grad_combine_kernel(CombineArgs q, int cv_log2) {
  pdl_sync();
  constexpr int VN = Vec<T>::N;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int xp = t >> cv_log2;
  if (xp >= (int)q.dst.wp) return;
  const int c = (t & ((1 << cv_log2) - 1)) * VN;
  const int yp = blockIdx.y, n = blockIdx.z;
  const int y = yp - q.dst.halo, x = xp - q.dst.halo;
  float v[VN];
#pragma unroll
  for (int k = 0; k < VN; ++k) v[k] = 0.f;
  const bool interior = (y >= 0 && y < q.dst.h && x >= 0 && x < q.dst.w);
  if (interior) {
    // every operand's 16 bytes are requested before the first one is used (one memory latency per thread, not one per operand)
    const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
    uint4 ra = z4, rb = z4, rc = z4, rm2 = z4, rmul = z4, rmask = z4;
    const T* ab = static_cast<const T*>(q.a.data);
    if (q.has_a) ra = ldg16(ab + toff(q.a, n, y + q.pa, x + q.pa, q.a_c_off + c));
    if (q.has_b) rb = ldg16(static_cast<const T*>(q.b.data) + toff(q.b, n, y, x, q.b_c_off + c));
    if (q.has_c) rc = ldg16(static_cast<const T*>(q.c.data) + toff(q.c, n, y, x, q.c_c_off + c));
    if (q.has_dst2) rm2 = ldg16(static_cast<const T*>(q.mul2.data) + toff(q.mul2, n, y, x, q.mul2_c_off + c));
    if (q.has_mul) rmul = ldg16(static_cast<const T*>(q.mul.data) + toff(q.mul, n, y, x, q.mul_c_off + c));
    if (q.has_mask) rmask = ldg16(static_cast<const T*>(q.mask.data) + toff(q.mask, n, y, x, q.mask_c_off + c));
    // per-tensor power-of-two scales: every addend is brought to the destination's scale (exact multiplications)
    const float so = tscale(q.dst);
    const float fa = q.has_a ? so * tinv(q.a) : 1.f, fb = q.has_b ? so * tinv(q.b) : 1.f, fc = q.has_c ? so * tinv(q.c) : 1.f;
    float t[VN];
    if (q.has_a) {
      cvt16<T>(ra, t);
#pragma unroll
      for (int k = 0; k < VN; ++k) v[k] += t[k] * fa;
      if (q.reflect_a && q.pa > 0) {
        // further positions of the padded tensor that map to (y, x): only within `pa` of an edge
        int ys[3], xs[3], ny = 0, nx = 0;
        ys[ny++] = y + q.pa;
        xs[nx++] = x + q.pa;
        if (y >= 1 && y <= q.pa) ys[ny++] = q.pa - y;
        if (y <= q.dst.h - 2 && y >= q.dst.h - 1 - q.pa) ys[ny++] = q.pa + 2 * (q.dst.h - 1) - y;
        if (x >= 1 && x <= q.pa) xs[nx++] = q.pa - x;
        if (x <= q.dst.w - 2 && x >= q.dst.w - 1 - q.pa) xs[nx++] = q.pa + 2 * (q.dst.w - 1) - x;
        if (ny * nx > 1)
          for (int iy = 0; iy < ny; ++iy)
            for (int ix = iy == 0 ? 1 : 0; ix < nx; ++ix) {
              Vec<T>::load(ab + toff(q.a, n, ys[iy], xs[ix], q.a_c_off + c), t);
#pragma unroll
              for (int k = 0; k < VN; ++k) v[k] += t[k] * fa;
            }
      }
    }
    if (q.has_b) {
      cvt16<T>(rb, t);
#pragma unroll
      for (int k = 0; k < VN; ++k) v[k] += t[k] * fb;
    }
    if (q.has_c) {
      cvt16<T>(rc, t);
#pragma unroll
      for (int k = 0; k < VN; ++k) v[k] += t[k] * fc;
    }
    if (q.has_dst2) {
      float o2[VN];
      const float f2 = tscale(q.dst2) * tinv(q.mul2) * pow2_rcp(so);  // the sum is in dst's scale: powers of two, exact
      cvt16<T>(rm2, t);
#pragma unroll
      for (int k = 0; k < VN; ++k) o2[k] = v[k] * (t[k] * f2);
      Vec<T>::store(static_cast<T*>(q.dst2.data) + toff(q.dst2, n, y, x, q.dst2_c_off + c), o2);
    }
    if (q.has_mul) {
      const float fm = tinv(q.mul);
      cvt16<T>(rmul, t);
#pragma unroll
      for (int k = 0; k < VN; ++k) v[k] *= t[k] * fm;
    }
    if (q.has_mask) {
      cvt16<T>(rmask, t);
#pragma unroll
      for (int k = 0; k < VN; ++k) {
        if (q.act == UEGAN_ACT_LRELU) v[k] *= (t[k] > 0.f ? 1.f : 0.2f);
        else if (q.act == UEGAN_ACT_RELU) v[k] = t[k] > 0.f ? v[k] : 0.f;
      }
    }
  }
  Vec<T>::store(static_cast<T*>(q.dst.data) + toff(q.dst, n, y, x, q.dst_c_off + c), v);
}

// ------------------------------------------------------------------------------------------
// fold_inplace: t (halo = pad) holds the gradient w.r.t. the reflect-PADDED input of a conv (what a dgrad launch writes:
// extent (h + 2 pad) x (w + 2 pad)).  Adjoint of nn.ReflectionPad2d in place: every interior pixel within `pad` of an
// edge adds the halo pixels that were reflected copies of it and zeroes them (each has exactly one reader), so the
// tensor is at once the folded gradient and a zero-haloed dgrad / wgrad operand.  Sources are halo pixels only, targets
// interior pixels only: race-free.  Touches O(perimeter) data instead of a full read + write pass (grad_combine).
// grid = (1, h, n); a row inside the top / bottom band processes every column, any other row only its 2*pad edge columns.
// ------------------------------------------------------------------------------------------
// grid = (items of one image / 256, images): one thread per (receiving pixel, 16-byte channel vector).  The receiving
// pixels are the `band` rows 1 .. pa and h-1-pa .. h-2 in full, plus the 2*pa edge columns of every other row (narrow
// images: every pixel).  (One block per image row with a serial loop over its pixels -- the first version -- spent 46-125 us
// per launch on 16 K nearly empty blocks and four dependent memory latencies per item, r4z.)
template <typename T>
__global__ void fold_inplace_kernel(TGeom t, int cv_log2, int all, int items_band, int items) {
  pdl_sync();
  constexpr int VN = Vec<T>::N;
  const int it = blockIdx.x * blockDim.x + threadIdx.x;
  if (it >= items) return;
  const int n = blockIdx.y, pa = t.halo;
  const int c = (it & ((1 << cv_log2) - 1)) * VN;
  int y, x;
  if (all) {
    const int pix = it >> cv_log2;
    y = pix / t.w;
    x = pix - y * t.w;
  } else if (it < items_band) {
    const int pix = it >> cv_log2;
    const int j = pix / t.w;  // band row index: 0 .. 2*pa-1
    x = pix - j * t.w;
    y = j < pa ? j + 1 : t.h - 1 - 2 * pa + j;
  } else {
    const int pix = (it - items_band) >> cv_log2;
    const int i = pix / (2 * pa), xi = pix - i * (2 * pa);  // i: 0 .. h-2*pa-1 over the rows outside the bands
    y = i == 0 ? 0 : (i == t.h - 2 * pa - 1 ? t.h - 1 : pa + i);
    // edge columns that receive reflected copies: 1 .. pa and w-1-pa .. w-2
    x = xi < pa ? xi + 1 : t.w - 1 - 2 * pa + xi;
  }
  // interior coordinates of the reflected sources (they lie in the halo); index 0 of each axis is the pixel's own row / column
  const bool vy1 = y >= 1 && y <= pa, vy2 = y <= t.h - 2 && y >= t.h - 1 - pa;
  const bool vx1 = x >= 1 && x <= pa, vx2 = x <= t.w - 2 && x >= t.w - 1 - pa;
  if (!(vy1 || vy2 || vx1 || vx2)) return;
  const int yc[3] = {y, -y, 2 * (t.h - 1) - y}, xc[3] = {x, -x, 2 * (t.w - 1) - x};
  const bool vy[3] = {true, vy1, vy2}, vx[3] = {true, vx1, vx2};
  T* base = static_cast<T*>(t.data);
  T* dst = base + toff(t, n, y, x, c);
  // every halo pixel is the reflected copy of exactly ONE interior pixel, so it is read by exactly one thread: that thread
  // zeroes it (the tensor leaves as a zero-haloed operand without a separate halo pass).  All loads first, then the stores;
  // the eight candidate sources are static slots (registers, predicated), in grad_combine's summation order: (y, x)
  // first, then rows ascending, columns ascending.
  T* src[8];
  uint4 raw[8];
  bool ok[8];
#pragma unroll
  for (int iy = 0; iy < 3; ++iy)
#pragma unroll
    for (int ix = 0; ix < 3; ++ix) {
      if (iy == 0 && ix == 0) continue;
      const int sl = iy * 3 + ix - 1;
      ok[sl] = vy[iy] && vx[ix];
      src[sl] = base + toff(t, n, yc[iy], xc[ix], c);
    }
  const uint4 rd = *reinterpret_cast<const uint4*>(dst);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    raw[i] = make_uint4(0u, 0u, 0u, 0u);
    if (ok[i]) raw[i] = *reinterpret_cast<const uint4*>(src[i]);
  }
  float v[VN];
  cvt16<T>(rd, v);
#pragma unroll
  for (int i = 0; i < 8; ++i)
    if (ok[i]) {
      float sv[VN];
      cvt16<T>(raw[i], sv);
#pragma unroll
      for (int k = 0; k < VN; ++k) v[k] += sv[k];
      *reinterpret_cast<uint4*>(src[i]) = make_uint4(0u, 0u, 0u, 0u);
    }
  Vec<T>::store(dst, v);
}

// ------------------------------------------------------------------------------------------
// dz_hstack: E[n, y, q, s*cout + o] = dz[n, y, q - s, o] for s < k, q in [0, w + k - 1)   (zero where q - s leaves the
// image: dz carries a zero halo >= k - 1).  The horizontally unrolled output gradient of a tiny-Cout conv: with it the
// weight gradient dW[o][c][r][s] = sum_{y,q} xpad[y + r][q][c] * E[y][q][(s, o)] is the wgrad of a k x 1 convolution with
// k*cout output channels -- k times fewer MMAs than one accumulator per (r, s) tap (conv_wgrad.cu, vertical patch mode).
// ------------------------------------------------------------------------------------------
template <typename T, int COUT>
__global__ void dz_hstack_kernel(TGeom dz, TGeom e, int k, long long total) {
  pdl_sync();
  constexpr int VN = Vec<T>::N;  // dz stores one 16-byte vector per pixel (4 fp32 / 8 fp16 channels, COUT real ones)
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int q = (int)(i % e.w);
  const int y = (int)((i / e.w) % e.h);
  const int n = (int)(i / ((long long)e.w * e.h));
  float v[32];
#pragma unroll
  for (int j = 0; j < 32; ++j) v[j] = 0.f;
  const T* db = static_cast<const T*>(dz.data);
#pragma unroll
  for (int s = 0; s < 7; ++s) {
    if (s < k) {
      float d[VN];
      Vec<T>::load(db + toff(dz, n, y, q - s, 0), d);
      v[s * COUT] = d[0];
      if (COUT > 1) v[s * COUT + 1] = d[1];
      if (COUT > 2) v[s * COUT + 2] = d[2];
    }
  }
  // (the stack shares dz's scale: a pure copy; Vec<float>::store re-rounds tf32 values to themselves)
  T* ep = static_cast<T*>(e.data) + toff(e, n, y, q, 0);
#pragma unroll
  for (int j = 0; j < 32; j += VN) {
    float o[VN];
#pragma unroll
    for (int t = 0; t < VN; ++t) o[t] = v[j + t];
    Vec<T>::store(ep + j, o);
  }
}

// ------------------------------------------------------------------------------------------
// per-channel sum over n, h, w (bias gradient): out[c] += sum
// ------------------------------------------------------------------------------------------
template <typename T>
struct ChannelSumOp {
  TGeom s;
  int c_off;
  float* out;
  int out_channels;
  __device__ void prep(int, int) {}
  __device__ void acc(int n, int y, int x, int c, float (&a)[Vec<T>::N][1]) const {
    float v[Vec<T>::N];
    Vec<T>::load(static_cast<const T*>(s.data) + toff(s, n, y, x, c_off + c), v);
#pragma unroll
    for (int k = 0; k < Vec<T>::N; ++k) a[k][0] += v[k];
  }
  __device__ void flush(int n, int c, const float (&t)[1]) const {
    if (c < out_channels) atomicAdd(out + c, t[0] * tinv(s));  // the bias gradient is a true-scale fp32 value
  }
};

// ------------------------------------------------------------------------------------------
// InstanceNorm backward.  xhat = (z - mean) * rstd ; out = xhat
//   dz = rstd * (dout - mean_p(dout) - xhat * mean_p(dout * xhat))
// ------------------------------------------------------------------------------------------
template <typename T>
struct InBwdStatsOp {
  TGeom g; int d_c_off;
  TGeom z;
  const float* mr;
  double* sums;
  float mean[Vec<T>::N], rstd[Vec<T>::N];
  __device__ void prep(int n, int c) {
#pragma unroll
    for (int k = 0; k < Vec<T>::N; ++k) {
      const long long si = ((long long)n * z.c + c + k) * 2;
      mean[k] = mr[si]; rstd[k] = mr[si + 1];
    }
  }
  __device__ void acc(int n, int y, int x, int c, float (&a)[Vec<T>::N][2]) const {
    float gv[Vec<T>::N], zv[Vec<T>::N];
    Vec<T>::load(static_cast<const T*>(g.data) + toff(g, n, y, x, d_c_off + c), gv);
    Vec<T>::load(static_cast<const T*>(z.data) + toff(z, n, y, x, c), zv);
#pragma unroll
    for (int k = 0; k < Vec<T>::N; ++k) {
      const float xh = (zv[k] - mean[k]) * rstd[k];
      a[k][0] += gv[k];
      a[k][1] += gv[k] * xh;
    }
  }
  __device__ void flush(int n, int c, const float (&t)[2]) const {
    atomicAdd(&sums[((long long)n * z.c + c) * 2], (double)t[0]);
    atomicAdd(&sums[((long long)n * z.c + c) * 2 + 1], (double)t[1]);
  }
};
// The "apply" pass of a normalisation backward needs six per-(n, c) statistics per element.  The block stages them for
// its image in shared memory (prologue: one thread per channel), then streams ROWS padded rows of one x-chunk:
// grid = (x-chunks, row groups, images), no per-element divisions, no per-element global statistics loads.  The
// per-element arithmetic is unchanged (bit-identical to evaluating the formulas below from global memory):
//   mode 0  InstanceNorm backward: a = dout, b = z:  xhat = (b - mean)*rstd ; dz = rstd*(a - m1 - xhat*m2)
//   mode 1  VGG tap (InstanceNorm + MSE) backward: a = x, b = y:  xh, yh normalised ; e = xh - yh ;
//           g = (cf*rstd_x)*(e - m1 - xh*m2) (+ deep), then the ReLU mask of x
// with m1 = mean_p(g), m2 = mean_p(g * xhat) from the strip-reduce pass before.
struct AffineArgs {
  TGeom a; int a_c_off;
  TGeom b;
  TGeom deep; int has_deep;
  TGeom dst;
  const float* mra;   // (mean, rstd) pairs of the normalised tensor (mode 0: of b = z; mode 1: of a = x)
  const float* mrb;   // mode 1: (mean, rstd) of y
  const double* sums; // per-(n, c) { sum(g), sum(g * xhat) }
  double inv_npix;
  float coef;
  const float* gscale;
  const float *s_num0, *s_num1, *s_den;  // optional per-tensor scales folded into the coefficient: * s_num0 * s_num1 / s_den
  int mode, cch, cv_log2;
};
// row groups per block: the statistics prologue is paid once per ROWS * kAffineIters rows (fp32: 4 trips measured faster,
// 1.08 -> 0.96 ms per step; the 16-bit tap backward slower, 1.55 -> 1.79 ms, so it keeps one trip)
template <typename T, typename TG, int ROWS>
__global__ void __launch_bounds__(256) affine_apply_kernel(const AffineArgs q) {
  pdl_sync();
  constexpr int kAffineIters = sizeof(T) == 4 ? 4 : 1;
  extern __shared__ float s_coef[];  // [6][cch]: mean_a, rstd_a (or cf*rstd_x), mean_b, rstd_b, m1, m2
  constexpr int VN = Vec<T>::N;
  const int n = blockIdx.z, C = q.cch;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int xp = t >> q.cv_log2;
  const bool live = xp < (int)q.dst.wp;
  const int c = (t & ((1 << q.cv_log2) - 1)) * VN;
  const int x = xp - q.dst.halo;
  const bool xin = live && x >= 0 && x < q.dst.w;
  for (int it = 0; it < kAffineIters; ++it) {
  const int row0 = ((int)blockIdx.y * kAffineIters + it) * ROWS;  // first padded row of this trip
  if (row0 >= (int)q.dst.hp) break;                               // (block-uniform; never on the first trip)
  // the data loads of the trip go out first; the statistics prologue of the block then overlaps their latency
  const uint4 z4 = make_uint4(0u, 0u, 0u, 0u);
  uint4 ar[ROWS], br[ROWS], dr[ROWS];
  bool in[ROWS];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int y = row0 + r - q.dst.halo;
    in[r] = xin && y >= 0 && y < q.dst.h;
    ar[r] = br[r] = dr[r] = z4;
    if (in[r]) {
      ar[r] = ldg16(static_cast<const T*>(q.a.data) + toff(q.a, n, y, x, q.a_c_off + c));
      br[r] = ldg16(static_cast<const T*>(q.b.data) + toff(q.b, n, y, x, c));
      if (q.has_deep) dr[r] = ldg16(static_cast<const TG*>(q.deep.data) + toff(q.deep, n, y, x, c));
    }
  }
  if (it == 0) {
    const float cf = q.coef * (q.gscale ? __ldg(q.gscale) : 1.f) * (q.s_num0 ? __ldg(q.s_num0) : 1.f) *
                     (q.s_num1 ? __ldg(q.s_num1) : 1.f) / (q.s_den ? __ldg(q.s_den) : 1.f);
    for (int ch = threadIdx.x; ch < C; ch += blockDim.x) {
      const long long si = ((long long)n * C + ch) * 2;
      s_coef[ch] = q.mra[si];
      s_coef[C + ch] = q.mra[si + 1];
      s_coef[2 * C + ch] = q.mode == 1 ? q.mrb[si] : 0.f;
      s_coef[3 * C + ch] = q.mode == 1 ? q.mrb[si + 1] : 0.f;
      s_coef[4 * C + ch] = (float)(q.sums[si] * q.inv_npix);
      s_coef[5 * C + ch] = (float)(q.sums[si + 1] * q.inv_npix);
    }
    if (threadIdx.x == 0) s_coef[6 * C] = cf;
    __syncthreads();
  }
  if (!live) continue;
  const float cf = s_coef[6 * C];
  float av[ROWS][VN], bv[ROWS][VN], dv[ROWS][VN];
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    cvt16<T>(ar[r], av[r]);
    cvt16<T>(br[r], bv[r]);
    cvt16<TG>(dr[r], dv[r]);
  }
  // (out-of-image positions hold zeros and are computed like any other, then selected away: no per-element branches)
  float v[ROWS][VN];
  if (q.mode == 0) {
    // stored tensors: xh is the true normalised value (ra = rstd / s_z), the bracket is in dout's stored units;
    // s_coef[6 * C] (= cf) carries s_z * s_dz / s_dout, which turns ra * bracket into dz's stored units
#pragma unroll
    for (int k = 0; k < VN; ++k) {
      const float ma = s_coef[c + k], ra = s_coef[C + c + k], m1 = s_coef[4 * C + c + k], m2 = s_coef[5 * C + c + k];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const float xh = (bv[r][k] - ma) * ra;
        const float g = cf * ra * (av[r][k] - m1 - xh * m2);
        v[r][k] = in[r] ? g : 0.f;
      }
    }
  } else {
    const bool deep = q.has_deep != 0;
#pragma unroll
    for (int k = 0; k < VN; ++k) {
      const float ma = s_coef[c + k], ra = s_coef[C + c + k], mb = s_coef[2 * C + c + k], rb = s_coef[3 * C + c + k];
      const float m1 = s_coef[4 * C + c + k], m2 = s_coef[5 * C + c + k];
#pragma unroll
      for (int r = 0; r < ROWS; ++r) {
        const float xh = (av[r][k] - ma) * ra;
        const float yh = (bv[r][k] - mb) * rb;
        const float e = xh - yh;
        const float g0 = cf * ra * (e - m1 - xh * m2);
        const float g = deep ? g0 + dv[r][k] : g0;
        v[r][k] = (in[r] && av[r][k] > 0.f) ? g : 0.f;  // ReLU mask of the tap
      }
    }
  }
#pragma unroll
  for (int r = 0; r < ROWS; ++r) {
    const int yp = row0 + r;
    if (yp >= (int)q.dst.hp) break;
    Vec<TG>::store(static_cast<TG*>(q.dst.data) + toff(q.dst, n, yp - q.dst.halo, x, c), v[r]);
  }
  }
}
template <typename T, typename TG>
static int launch_affine_apply(AffineArgs& q, cudaStream_t st) {
  constexpr int VN = Vec<T>::N, ROWS = 2;
  const int cv = q.cch / VN;
  int lg = 0;
  while ((1 << lg) < cv) ++lg;
  UEGAN_CHECK((1 << lg) == cv, "normalisation backward: channels / vector width must be a power of two (got %d)", cv);
  constexpr int RB = ROWS * (sizeof(T) == 4 ? 4 : 1);  // padded rows per block (kAffineIters of the kernel)
  UEGAN_CHECK(q.dst.n <= 65535 && (q.dst.hp + RB - 1) / RB <= 65535, "normalisation backward: tensor too large");
  q.cv_log2 = lg;
  const dim3 grid((unsigned)((q.dst.wp * cv + 255) / 256), (unsigned)((q.dst.hp + RB - 1) / RB), (unsigned)q.dst.n);
  launch_pdl(affine_apply_kernel<T, TG, ROWS>, grid, 256, (6 * q.cch + 1) * sizeof(float), st, q);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

// ------------------------------------------------------------------------------------------
// bilinear x2 (align_corners) backward, gather form: exact adjoint of upsample2x_kernel
// ------------------------------------------------------------------------------------------
// For the x2 case only output rows 2*yi-2 .. 2*yi+3 can have floor(sy*yo) in {yi-1, yi} (1/sy = 2 + 1/(h-1)); the
// weights are derived from the SAME float expression the forward kernel evaluates, so the pair is an exact adjoint.
// grid = (x-chunks of an input row, input rows, images).
template <typename T>
__global__ void upsample2x_bwd_kernel(TGeom g, int g_c_off, TGeom d, float sy, float sx, int cv_log2) {
  pdl_sync();
  constexpr int VN = Vec<T>::N;
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int xi = t >> cv_log2;
  if (xi >= d.w) return;
  const int c = (t & ((1 << cv_log2) - 1)) * VN;
  const int yi = blockIdx.y, n = blockIdx.z;
  float wy[6], wx[6];
#pragma unroll
  for (int j = 0; j < 6; ++j) {
    const int yo = 2 * yi - 2 + j;
    float w = 0.f;
    if (yo >= 0 && yo < g.h) {
      const float fy = sy * yo;
      const int y0 = (int)fy;
      const int y1 = y0 + (y0 < d.h - 1 ? 1 : 0);
      const float ly = fy - y0;
      if (y0 == yi) w += 1.f - ly;
      if (y1 == yi) w += ly;
    }
    wy[j] = w;
    const int xo = 2 * xi - 2 + j;
    w = 0.f;
    if (xo >= 0 && xo < g.w) {
      const float fx = sx * xo;
      const int x0 = (int)fx;
      const int x1 = x0 + (x0 < d.w - 1 ? 1 : 0);
      const float lx = fx - x0;
      if (x0 == xi) w += 1.f - lx;
      if (x1 == xi) w += lx;
    }
    wx[j] = w;
  }
  float acc[VN];
#pragma unroll
  for (int k = 0; k < VN; ++k) acc[k] = 0.f;
  const T* gb = static_cast<const T*>(g.data);
  // (requesting two candidate rows of taps at once with predicated loads was measured 33 % slower, r4z: 78 registers)
#pragma unroll
  for (int jy = 0; jy < 6; ++jy) {
    if (wy[jy] == 0.f) continue;
    const T* gr = gb + toff(g, n, 2 * yi - 2 + jy, 2 * xi - 2, g_c_off + c);
#pragma unroll
    for (int jx = 0; jx < 6; ++jx) {
      if (wx[jx] == 0.f) continue;
      float tv[VN];
      Vec<T>::load(gr + (long long)jx * g.c, tv);
#pragma unroll
      for (int k = 0; k < VN; ++k) acc[k] += wy[jy] * wx[jx] * tv[k];  // same order as the scatter's adjoint sum
    }
  }
  const float rs = tscale(d) * tinv(g);
#pragma unroll
  for (int k = 0; k < VN; ++k) acc[k] *= rs;
  Vec<T>::store(static_cast<T*>(d.data) + toff(d, n, yi, xi, c), acc);
}

// ------------------------------------------------------------------------------------------
// max-pool 2x2 backward fused with the ReLU mask of the pooled conv's output and an optional extra gradient:
//   dsrc[2yo+a, 2xo+b] = (src is the argmax of its window ? dpool[yo, xo] : 0)   (first max wins, like PyTorch)
// written with a zero halo.
// ------------------------------------------------------------------------------------------
template <typename T, typename TG>
__global__ void maxpool2x2_bwd_kernel(TGeom src, TGeom dpool, TGeom dst, long long total) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  constexpr int VN = 8;
  const int cv = src.c / VN;
  const int c = (int)(i % cv) * VN;
  long long pix = i / cv;
  const int xo = (int)(pix % dpool.w);
  pix /= dpool.w;
  const int yo = (int)(pix % dpool.h);
  const int n = (int)(pix / dpool.h);
  float v[4][VN], g[VN];
  const T* sb = static_cast<const T*>(src.data);
  Vec<T>::load(sb + toff(src, n, 2 * yo, 2 * xo, c), v[0]);
  Vec<T>::load(sb + toff(src, n, 2 * yo, 2 * xo + 1, c), v[1]);
  Vec<T>::load(sb + toff(src, n, 2 * yo + 1, 2 * xo, c), v[2]);
  Vec<T>::load(sb + toff(src, n, 2 * yo + 1, 2 * xo + 1, c), v[3]);
  Vec<TG>::load(static_cast<const TG*>(dpool.data) + toff(dpool, n, yo, xo, c), g);
  float o[4][VN];
#pragma unroll
  for (int k = 0; k < VN; ++k) {
    int best = 0;
    float bv = v[0][k];
#pragma unroll
    for (int j = 1; j < 4; ++j)
      if (v[j][k] > bv) { bv = v[j][k]; best = j; }
#pragma unroll
    for (int j = 0; j < 4; ++j) o[j][k] = (j == best && bv > 0.f) ? g[k] : 0.f;  // + ReLU mask of the activation
  }
  TG* db = static_cast<TG*>(dst.data);
  Vec<TG>::store(db + toff(dst, n, 2 * yo, 2 * xo, c), o[0]);
  Vec<TG>::store(db + toff(dst, n, 2 * yo, 2 * xo + 1, c), o[1]);
  Vec<TG>::store(db + toff(dst, n, 2 * yo + 1, 2 * xo, c), o[2]);
  Vec<TG>::store(db + toff(dst, n, 2 * yo + 1, 2 * xo + 1, c), o[3]);
}

// ------------------------------------------------------------------------------------------
// VGG tap: d/dx of  weight * mean((IN(x) - IN(y))^2)  w.r.t. the feature map x (y is a constant):
//   e = xh - yh ; dxh = (2 * weight / numel) * e ; dx = rstd * (dxh - mean_p(dxh) - xh * mean_p(dxh * xh))
// followed by the ReLU mask of the tap itself (x is a ReLU output) and accumulation with the gradient that arrives from
// deeper layers (`deep`, optional).  Output gradient tensor TG (fp16, loss-scaled by the caller through `weight`:
// per-element gradients of a mean over 1e8 elements are far below fp16's range unscaled), zero halo.
// ------------------------------------------------------------------------------------------
template <typename T>
struct TapBwdStatsOp {
  TGeom x, y;
  const float* mrx;
  const float* mry;
  double* sums;
  float mx[Vec<T>::N], rx[Vec<T>::N], my[Vec<T>::N], ry[Vec<T>::N];
  __device__ void prep(int n, int c) {
#pragma unroll
    for (int k = 0; k < Vec<T>::N; ++k) {
      const long long si = ((long long)n * x.c + c + k) * 2;
      mx[k] = mrx[si]; rx[k] = mrx[si + 1]; my[k] = mry[si]; ry[k] = mry[si + 1];
    }
  }
  __device__ void acc(int n, int yy, int xx, int c, float (&a)[Vec<T>::N][2]) const {
    float xv[Vec<T>::N], yv[Vec<T>::N];
    Vec<T>::load(static_cast<const T*>(x.data) + toff(x, n, yy, xx, c), xv);
    Vec<T>::load(static_cast<const T*>(y.data) + toff(y, n, yy, xx, c), yv);
#pragma unroll
    for (int k = 0; k < Vec<T>::N; ++k) {
      const float xh = (xv[k] - mx[k]) * rx[k];
      const float e = xh - (yv[k] - my[k]) * ry[k];
      a[k][0] += e;
      a[k][1] += e * xh;
    }
  }
  __device__ void flush(int n, int c, const float (&t)[2]) const {
    atomicAdd(&sums[((long long)n * x.c + c) * 2], (double)t[0]);
    atomicAdd(&sums[((long long)n * x.c + c) * 2 + 1], (double)t[1]);
  }
};
// gradient of the packed input: NHWC (c >= 3) -> NCHW fp32 (3 channels) times per-channel scale
template <typename TG>
__global__ void unpack_grad_kernel(TGeom s, float* __restrict__ dst, float s0, float s1, float s2, long long total,
                                   const float* __restrict__ skip_dout, const float* __restrict__ skip_res,
                                   const float* __restrict__ skip_x) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int x = (int)(i % s.w);
  long long r = i / s.w;
  const int y = (int)(r % s.h);
  r /= s.h;
  const int c = (int)(r % 3);
  const int n = (int)(r / 3);
  const float sc = c == 0 ? s0 : (c == 1 ? s1 : s2);
  float v = to_f32<TG>(static_cast<const TG*>(s.data)[toff(s, n, y, x, c)]) * sc * tinv(s);
  if (skip_dout) {
    // the generator's identity path out = clamp(res + x, -1, 1) (models.py:72): d out / d x = [|res + x| <= 1]
    const float t = __ldg(skip_res + i) + __ldg(skip_x + i);
    if (t >= -1.f && t <= 1.f) v += __ldg(skip_dout + i);
  }
  dst[i] = v;
}

static inline unsigned nblk(long long total, int threads) { return (unsigned)((total + threads - 1) / threads); }

}  // namespace uegan

using namespace uegan;

#define UEGAN_DISPATCH(dtype, KERNEL, ...)                      \
  do {                                                          \
    if ((dtype) == UEGAN_F32) launch_pdl(KERNEL<float>, __VA_ARGS__);        \
    else if ((dtype) == UEGAN_BF16) launch_pdl(KERNEL<__nv_bfloat16>, __VA_ARGS__); \
    else launch_pdl(KERNEL<__half>, __VA_ARGS__);                            \
  } while (0)

extern "C" {

int uegan_head_bwd(const float* dout_nchw, const float* out_nchw, const float* x_nchw, int32_t channels, int32_t mode,
                   const uegan_tensor* dz, void* stream) {
  UEGAN_CHECK(dout_nchw && out_nchw && dz && dz->data, "head_bwd: null pointer");
  UEGAN_CHECK(mode >= 0 && mode <= 2 && (mode != 2 || x_nchw), "head_bwd: bad mode");
  UEGAN_CHECK(dtype_ok(dz->dtype) && (dz->c * dtype_size(dz->dtype)) % 16 == 0, "head_bwd: bad dz tensor");
  UEGAN_CHECK(channels <= 16 / dtype_size(dz->dtype), "head_bwd: too many channels");
  const TGeom d = geom(*dz);
  const long long total = (long long)d.n * d.hp * d.wp;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  UEGAN_DISPATCH(dz->dtype, head_bwd_kernel, nblk(total, 256), 256, 0, st, dout_nchw, out_nchw, x_nchw, d, channels,
                                                                                mode, total);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

static int grad_combine_impl(const uegan_tensor* dst, int32_t dst_c_off, int32_t channels, const uegan_tensor* src_a,
                             int32_t a_c_off, int32_t pad_a, int32_t pad_mode_a, const uegan_tensor* add_b, int32_t b_c_off,
                             const uegan_tensor* add_c, int32_t c_c_off, const uegan_tensor* mask, int32_t mask_c_off,
                             int32_t act, const uegan_tensor* mul, int32_t mul_c_off, const uegan_tensor* dst2,
                             int32_t dst2_c_off, const uegan_tensor* mul2, int32_t mul2_c_off, void* stream) {
  UEGAN_CHECK(dst && dst->data && (src_a || add_b || add_c), "grad_combine: null pointer");
  CombineArgs q;
  memset(&q, 0, sizeof(q));
  q.dst = geom(*dst); q.dst_c_off = dst_c_off; q.cch = channels; q.act = act;
  const int vn = 16 / dtype_size(dst->dtype);
  UEGAN_CHECK(channels % vn == 0 && dst_c_off % vn == 0 && dst_c_off + channels <= dst->c, "grad_combine: bad channels");
  auto chk = [&](const uegan_tensor* t, int off, int extra, const char* who) -> int {
    UEGAN_CHECK(t->data && t->dtype == dst->dtype && t->n == dst->n && t->h == dst->h + 2 * extra &&
                    t->w == dst->w + 2 * extra && off % vn == 0 && off + channels <= t->c,
                "grad_combine: %s mismatch (%dx%dx%dx%d vs dst %dx%dx%dx%d, extra %d)", who, t->n, t->h, t->w, t->c,
                dst->n, dst->h, dst->w, dst->c, extra);
    return 0;
  };
  if (src_a) {
    if (chk(src_a, a_c_off, pad_a, "src_a")) return -1;
    q.a = geom(*src_a); q.a_c_off = a_c_off; q.pa = pad_a; q.reflect_a = pad_mode_a == UEGAN_PAD_REFLECT; q.has_a = 1;
    UEGAN_CHECK(!q.reflect_a || (pad_a < dst->h && pad_a < dst->w), "grad_combine: pad too large");
  }
  if (add_b) { if (chk(add_b, b_c_off, 0, "add_b")) return -1; q.b = geom(*add_b); q.b_c_off = b_c_off; q.has_b = 1; }
  if (add_c) { if (chk(add_c, c_c_off, 0, "add_c")) return -1; q.c = geom(*add_c); q.c_c_off = c_c_off; q.has_c = 1; }
  if (mask) { if (chk(mask, mask_c_off, 0, "mask")) return -1; q.mask = geom(*mask); q.mask_c_off = mask_c_off; q.has_mask = 1; }
  if (mul) { if (chk(mul, mul_c_off, 0, "mul")) return -1; q.mul = geom(*mul); q.mul_c_off = mul_c_off; q.has_mul = 1; }
  if (dst2) {
    UEGAN_CHECK(mul2 && dst2->halo == 0, "grad_combine2: the second output needs its factor mul2 and a halo of 0");
    if (chk(dst2, dst2_c_off, 0, "dst2") || chk(mul2, mul2_c_off, 0, "mul2")) return -1;
    q.dst2 = geom(*dst2); q.dst2_c_off = dst2_c_off; q.has_dst2 = 1;
    q.mul2 = geom(*mul2); q.mul2_c_off = mul2_c_off;
  }
  const int cv = channels / vn;
  int cv_log2 = 0;
  while ((1 << cv_log2) < cv) ++cv_log2;
  UEGAN_CHECK((1 << cv_log2) == cv, "grad_combine: channels / vector width must be a power of two (got %d)", cv);
  UEGAN_CHECK(q.dst.hp <= 65535 && q.dst.n <= 65535, "grad_combine: tensor too large for the launch grid");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid(nblk((long long)q.dst.wp * cv, 256), (unsigned)q.dst.hp, (unsigned)q.dst.n);
  UEGAN_DISPATCH(dst->dtype, grad_combine_kernel, grid, 256, 0, st, q, cv_log2);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_grad_combine(const uegan_tensor* dst, int32_t dst_c_off, int32_t channels, const uegan_tensor* src_a,
                       int32_t a_c_off, int32_t pad_a, int32_t pad_mode_a, const uegan_tensor* add_b, int32_t b_c_off,
                       const uegan_tensor* add_c, int32_t c_c_off, const uegan_tensor* mask, int32_t mask_c_off,
                       int32_t act, const uegan_tensor* mul, int32_t mul_c_off, void* stream) {
  return grad_combine_impl(dst, dst_c_off, channels, src_a, a_c_off, pad_a, pad_mode_a, add_b, b_c_off, add_c, c_c_off, mask,
                           mask_c_off, act, mul, mul_c_off, nullptr, 0, nullptr, 0, stream);
}

int uegan_grad_combine2(const uegan_tensor* dst, int32_t dst_c_off, int32_t channels, const uegan_tensor* src_a,
                        int32_t a_c_off, int32_t pad_a, int32_t pad_mode_a, const uegan_tensor* add_b, int32_t b_c_off,
                        const uegan_tensor* add_c, int32_t c_c_off, const uegan_tensor* mask, int32_t mask_c_off,
                        int32_t act, const uegan_tensor* mul, int32_t mul_c_off, const uegan_tensor* dst2,
                        int32_t dst2_c_off, const uegan_tensor* mul2, int32_t mul2_c_off, void* stream) {
  UEGAN_CHECK(dst2 && mul2, "grad_combine2: null second output");
  return grad_combine_impl(dst, dst_c_off, channels, src_a, a_c_off, pad_a, pad_mode_a, add_b, b_c_off, add_c, c_c_off, mask,
                           mask_c_off, act, mul, mul_c_off, dst2, dst2_c_off, mul2, mul2_c_off, stream);
}

int uegan_fold_inplace(const uegan_tensor* t, void* stream) {
  UEGAN_CHECK(t && t->data, "fold_inplace: null pointer");
  UEGAN_CHECK(dtype_ok(t->dtype) && (t->c * dtype_size(t->dtype)) % 16 == 0, "fold_inplace: bad tensor");
  if (t->halo == 0) return 0;
  UEGAN_CHECK(t->halo < t->h && t->halo < t->w, "fold_inplace: halo %d needs h, w > halo (got %dx%d)", t->halo, t->h, t->w);
  const TGeom g = geom(*t);
  const int vn = 16 / dtype_size(t->dtype);
  const int cv = t->c / vn;
  int lg = 0;
  while ((1 << lg) < cv) ++lg;
  UEGAN_CHECK((1 << lg) == cv, "fold_inplace: channels / vector width must be a power of two (got %d)", cv);
  UEGAN_CHECK(g.n <= 65535 && (long long)g.h * g.w * cv < (1ll << 30), "fold_inplace: tensor too large for the launch grid");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int pa = g.halo;
  const int all = (g.h <= 2 * pa + 1 || g.w <= 2 * pa + 1) ? 1 : 0;  // narrow: the band / edge sets would overlap
  const int items_band = all ? 0 : 2 * pa * g.w * cv;
  const int items = all ? g.h * g.w * cv : items_band + (g.h - 2 * pa) * 2 * pa * cv;
  const dim3 grid((unsigned)((items + 255) / 256), (unsigned)g.n);
  UEGAN_DISPATCH(t->dtype, fold_inplace_kernel, grid, 256, 0, st, g, lg, all, items_band, items);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_dz_hstack(const uegan_tensor* dz, int32_t cout, int32_t k, const uegan_tensor* e, void* stream) {
  UEGAN_CHECK(dz && e && dz->data && e->data, "dz_hstack: null pointer");
  UEGAN_CHECK(dz->dtype == e->dtype && (dz->dtype == UEGAN_F32 || dz->dtype == UEGAN_F16) &&
                  dz->c * dtype_size(dz->dtype) == 16 && e->c == 32 && e->halo == 0,
              "dz_hstack: expects a one-vector (16-byte) dz and a 32-channel stack without halo, both fp32 or both fp16");
  UEGAN_CHECK((cout == 1 || cout == 3) && k >= 1 && k <= 7 && k * cout <= 21, "dz_hstack: unsupported cout %d / k %d", cout, k);
  UEGAN_CHECK(dz->halo >= k - 1 && e->n == dz->n && e->h == dz->h && e->w == dz->w + k - 1,
              "dz_hstack: dz needs a zero halo >= k - 1 and the stack must be %d x %d (got %d x %d)", dz->h, dz->w + k - 1,
              e->h, e->w);
  const TGeom d = geom(*dz), g = geom(*e);
  const long long total = (long long)g.n * g.h * g.w;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (dz->dtype == UEGAN_F32) {
    if (cout == 1) launch_pdl(dz_hstack_kernel<float, 1>, nblk(total, 256), 256, 0, st, d, g, k, total);
    else launch_pdl(dz_hstack_kernel<float, 3>, nblk(total, 256), 256, 0, st, d, g, k, total);
  } else {
    if (cout == 1) launch_pdl(dz_hstack_kernel<__half, 1>, nblk(total, 256), 256, 0, st, d, g, k, total);
    else launch_pdl(dz_hstack_kernel<__half, 3>, nblk(total, 256), 256, 0, st, d, g, k, total);
  }
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_channel_sum(const uegan_tensor* src, int32_t c_off, int32_t channels, float* out, int32_t out_channels,
                      int32_t accumulate, void* stream) {
  UEGAN_CHECK(src && src->data && out, "channel_sum: null pointer");
  UEGAN_CHECK(c_off >= 0 && c_off + channels <= src->c && channels <= 1024, "channel_sum: bad channel range");
  if (out_channels <= 0 || out_channels > channels) out_channels = channels;
  const TGeom s = geom(*src);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!accumulate) UEGAN_CUDA(cudaMemsetAsync(out, 0, sizeof(float) * out_channels, st));
  const int vn = 16 / dtype_size(src->dtype);
  UEGAN_CHECK(channels % vn == 0 && c_off % vn == 0 && 256 % (channels / vn) == 0,
              "channel_sum: unsupported channel count %d", channels);
  if (src->dtype == UEGAN_F32) {
    ChannelSumOp<float> op{s, c_off, out, out_channels};
    launch_strip_reduce<float, 1>(op, channels, s.n, s.h, s.w, st);
  } else if (src->dtype == UEGAN_BF16) {
    ChannelSumOp<__nv_bfloat16> op{s, c_off, out, out_channels};
    launch_strip_reduce<__nv_bfloat16, 1>(op, channels, s.n, s.h, s.w, st);
  } else {
    ChannelSumOp<__half> op{s, c_off, out, out_channels};
    launch_strip_reduce<__half, 1>(op, channels, s.n, s.h, s.w, st);
  }
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_instance_norm_bwd(const uegan_tensor* dout, int32_t d_c_off, const uegan_tensor* z, const float* mean_rstd,
                            const uegan_tensor* dz, double* ws, void* stream) {
  UEGAN_CHECK(dout && z && dz && mean_rstd && ws, "instance_norm_bwd: null pointer");
  UEGAN_CHECK(dout->dtype == z->dtype && z->dtype == dz->dtype && dout->n == z->n && dout->h == z->h && dout->w == z->w &&
                  dz->n == z->n && dz->h == z->h && dz->w == z->w && dz->c == z->c && d_c_off + z->c <= dout->c,
              "instance_norm_bwd: tensor mismatch");
  const TGeom g = geom(*dout), zz = geom(*z), d = geom(*dz);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nc = zz.n * zz.c;
  UEGAN_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * nc, st));
  const long long npix = (long long)zz.h * zz.w;
  UEGAN_CHECK(256 % (zz.c / (16 / dtype_size(z->dtype))) == 0, "instance_norm_bwd: unsupported channel count %d", zz.c);
  if (z->dtype == UEGAN_F32) {
    InBwdStatsOp<float> op{g, d_c_off, zz, mean_rstd, ws};
    launch_strip_reduce<float, 2>(op, zz.c, zz.n, zz.h, zz.w, st);
  } else if (z->dtype == UEGAN_BF16) {
    InBwdStatsOp<__nv_bfloat16> op{g, d_c_off, zz, mean_rstd, ws};
    launch_strip_reduce<__nv_bfloat16, 2>(op, zz.c, zz.n, zz.h, zz.w, st);
  } else {
    InBwdStatsOp<__half> op{g, d_c_off, zz, mean_rstd, ws};
    launch_strip_reduce<__half, 2>(op, zz.c, zz.n, zz.h, zz.w, st);
  }
  AffineArgs q;
  memset(&q, 0, sizeof(q));
  q.a = g; q.a_c_off = d_c_off; q.b = zz; q.deep = d; q.has_deep = 0; q.dst = d;
  q.mra = mean_rstd; q.mrb = nullptr; q.sums = ws; q.inv_npix = 1.0 / (double)npix; q.coef = 1.f; q.gscale = nullptr;
  q.mode = 0; q.cch = zz.c;
  q.s_num0 = zz.scale; q.s_num1 = d.scale; q.s_den = g.scale;
  if (z->dtype == UEGAN_F32) { if (launch_affine_apply<float, float>(q, st)) return -1; }
  else if (z->dtype == UEGAN_BF16) { if (launch_affine_apply<__nv_bfloat16, __nv_bfloat16>(q, st)) return -1; }
  else { if (launch_affine_apply<__half, __half>(q, st)) return -1; }
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_upsample2x_bwd(const uegan_tensor* dout, int32_t d_c_off, const uegan_tensor* dsrc, void* stream) {
  UEGAN_CHECK(dout && dsrc, "upsample2x_bwd: null pointer");
  UEGAN_CHECK(dout->dtype == dsrc->dtype && dout->n == dsrc->n && dout->h == 2 * dsrc->h && dout->w == 2 * dsrc->w &&
                  d_c_off + dsrc->c <= dout->c,
              "upsample2x_bwd: tensor mismatch");
  const TGeom g = geom(*dout), d = geom(*dsrc);
  const float sy = g.h > 1 ? (float)(d.h - 1) / (float)(g.h - 1) : 0.f;
  const float sx = g.w > 1 ? (float)(d.w - 1) / (float)(g.w - 1) : 0.f;
  const int vn = 16 / dtype_size(dsrc->dtype);
  const int cv = d.c / vn;
  int lg = 0;
  while ((1 << lg) < cv) ++lg;
  UEGAN_CHECK((1 << lg) == cv, "upsample2x_bwd: channels / vector width must be a power of two (got %d)", cv);
  UEGAN_CHECK(d.h <= 65535 && d.n <= 65535, "upsample2x_bwd: tensor too large for the launch grid");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const dim3 grid(nblk((long long)d.w * cv, 256), (unsigned)d.h, (unsigned)d.n);
  UEGAN_DISPATCH(dsrc->dtype, upsample2x_bwd_kernel, grid, 256, 0, st, g, d_c_off, d, sy, sx, lg);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_maxpool2x2_bwd(const uegan_tensor* src, const uegan_tensor* dpool, const uegan_tensor* dsrc, void* stream) {
  UEGAN_CHECK(src && dpool && dsrc, "maxpool2x2_bwd: null pointer");
  UEGAN_CHECK(src->dtype == UEGAN_F16 && dpool->dtype == UEGAN_F16 && dsrc->dtype == UEGAN_F16,
              "maxpool2x2_bwd: expects fp16 activations and (loss-scaled) fp16 gradients");
  UEGAN_CHECK(src->n == dsrc->n && src->h == dsrc->h && src->w == dsrc->w && src->c == dsrc->c && dpool->h == src->h / 2 &&
                  dpool->w == src->w / 2 && dpool->c == src->c && src->c % 8 == 0,
              "maxpool2x2_bwd: tensor mismatch");
  const TGeom s = geom(*src), g = geom(*dpool), d = geom(*dsrc);
  const long long total = (long long)g.n * g.h * g.w * (s.c / 8);
  launch_pdl(maxpool2x2_bwd_kernel<__half, __half>, nblk(total, 256), 256, 0, static_cast<cudaStream_t>(stream), s, g, d, total);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

static int in_mse_bwd_impl(const uegan_tensor* x, const uegan_tensor* y, const float* mean_rstd_x, const float* mean_rstd_y,
                           float weight, const float* gscale_dev, const uegan_tensor* deep, const uegan_tensor* dx, double* ws,
                           bool sums_ready, void* stream) {
  UEGAN_CHECK(x && y && dx && mean_rstd_x && mean_rstd_y && ws, "in_mse_bwd: null pointer");
  UEGAN_CHECK(x->dtype == UEGAN_F16 && y->dtype == UEGAN_F16 && dx->dtype == UEGAN_F16 && x->c % 8 == 0 && x->c <= 1024,
              "in_mse_bwd: expects fp16 features and a (loss-scaled) fp16 gradient");
  UEGAN_CHECK(x->n == y->n && x->h == y->h && x->w == y->w && x->c == y->c && dx->n == x->n && dx->h == x->h &&
                  dx->w == x->w && dx->c == x->c,
              "in_mse_bwd: tensor mismatch");
  if (deep) UEGAN_CHECK(deep->dtype == UEGAN_F16 && deep->h == x->h && deep->w == x->w && deep->c == x->c, "in_mse_bwd: deep mismatch");
  const TGeom gx = geom(*x), gy = geom(*y), gd = geom(*dx);
  TGeom gdeep = gd;
  if (deep) gdeep = geom(*deep);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const int nc = gx.n * gx.c;
  const long long npix = (long long)gx.h * gx.w;
  UEGAN_CHECK(256 % (gx.c / 8) == 0, "in_mse_bwd: unsupported channel count %d", gx.c);
  if (!sums_ready) {
    UEGAN_CUDA(cudaMemsetAsync(ws, 0, sizeof(double) * 2 * nc, st));
    TapBwdStatsOp<__half> op{gx, gy, mean_rstd_x, mean_rstd_y, ws};
    launch_strip_reduce<__half, 2>(op, gx.c, gx.n, gx.h, gx.w, st);
  }
  const double numel = (double)gx.n * gx.c * (double)npix;
  AffineArgs q;
  memset(&q, 0, sizeof(q));
  q.a = gx; q.a_c_off = 0; q.b = gy; q.deep = gdeep; q.has_deep = deep ? 1 : 0; q.dst = gd;
  q.mra = mean_rstd_x; q.mrb = mean_rstd_y; q.sums = ws; q.inv_npix = 1.0 / (double)npix;
  q.coef = (float)(2.0 * weight / numel); q.gscale = gscale_dev; q.mode = 1; q.cch = gx.c;
  if (launch_affine_apply<__half, __half>(q, st)) return -1;
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

int uegan_in_mse_bwd(const uegan_tensor* x, const uegan_tensor* y, const float* mean_rstd_x, const float* mean_rstd_y,
                     float weight, const float* gscale_dev, const uegan_tensor* deep, const uegan_tensor* dx, double* ws,
                     void* stream) {
  return in_mse_bwd_impl(x, y, mean_rstd_x, mean_rstd_y, weight, gscale_dev, deep, dx, ws, false, stream);
}

int uegan_in_mse_bwd_apply(const uegan_tensor* x, const uegan_tensor* y, const float* mean_rstd_x, const float* mean_rstd_y,
                           float weight, const float* gscale_dev, const uegan_tensor* deep, const uegan_tensor* dx,
                           const double* sums, void* stream) {
  return in_mse_bwd_impl(x, y, mean_rstd_x, mean_rstd_y, weight, gscale_dev, deep, dx, const_cast<double*>(sums), true, stream);
}

int uegan_unpack_input_grad(const uegan_tensor* dx, const float* scale_host, float* dst_nchw, const float* skip_dout_nchw,
                            const float* skip_res_nchw, const float* skip_x_nchw, void* stream) {
  UEGAN_CHECK(dx && dx->data && dst_nchw && dx->c >= 3, "unpack_input_grad: null pointer");
  UEGAN_CHECK(!skip_dout_nchw || (skip_res_nchw && skip_x_nchw), "unpack_input_grad: the identity path needs res and x");
  const TGeom s = geom(*dx);
  const long long total = (long long)s.n * 3 * s.h * s.w;
  const float s0 = scale_host ? scale_host[0] : 1.f, s1 = scale_host ? scale_host[1] : 1.f, s2 = scale_host ? scale_host[2] : 1.f;
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  UEGAN_DISPATCH(dx->dtype, unpack_grad_kernel, nblk(total, 256), 256, 0, st, s, dst_nchw, s0, s1, s2, total,
                                                                                   skip_dout_nchw, skip_res_nchw, skip_x_nchw);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
