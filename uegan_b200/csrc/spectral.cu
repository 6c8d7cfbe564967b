// Spectral normalisation of the Discriminator's strided convs (reference: models.py:185-188 ->
// torch.nn.utils.spectral_norm, n_power_iterations=1, dim=0, eps=1e-12).  W is the OIHW fp32 weight viewed as a
// (rows = cout) x (cols = cin*k*k) row-major matrix.
//   train:  v <- normalize(W^T u);  u <- normalize(W v);  sigma = u . (W v)
//   eval :  sigma = u . (W v) with the stored u, v
// The convolution consumes the UN-normalised packed weight; 1/sigma is applied as the epilogue's `alpha`.
#include "common.cuh"
#include "host_util.h"

namespace uegan {

// t[col] = sum_row W[row][col] * u[row];  threads along columns (coalesced), each block covers a row slab.
__global__ void sn_wt_u_kernel(const float* __restrict__ w, const float* __restrict__ u, float* __restrict__ t,
                               int rows, int cols, int rows_per_block) {
  pdl_sync();
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  if (col >= cols) return;
  const int r0 = blockIdx.y * rows_per_block;
  int r1 = r0 + rows_per_block;
  if (r1 > rows) r1 = rows;
  float acc = 0.f;
  for (int r = r0; r < r1; ++r) acc += __ldg(w + (long long)r * cols + col) * __ldg(u + r);
  atomicAdd(t + col, acc);
}

// nrm2[0] = sum_i t[i]^2
__global__ void sn_sumsq_kernel(const float* __restrict__ t, int n, double* __restrict__ out) {
  pdl_sync();
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc += (double)t[i] * t[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// optional v <- t / max(||t||, eps) (train), then wv[row] = W[row] . v ; one warp per row
__global__ void sn_w_v_kernel(const float* __restrict__ w, const float* __restrict__ t, const double* __restrict__ t_sumsq,
                              float* __restrict__ v, float* __restrict__ wv, int rows, int cols, int train, float eps) {
  pdl_sync();
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  float inv = 1.f;
  if (train) inv = 1.f / fmaxf((float)sqrt(*t_sumsq), eps);
  const float* src = train ? t : v;
  float acc = 0.f;
  for (int c = lane; c < cols; c += 32) {
    const float vv = src[c] * inv;
    acc += __ldg(w + (long long)warp * cols + c) * vv;
    if (train && warp == 0) v[c] = vv;  // row 0's warp also publishes the new v
  }
  acc = warp_sum(acc);
  if (lane == 0) wv[warp] = acc;
}

// single block: u <- normalize(wv) (train); sigma = u . wv; out[0] = sigma, out[1] = 1/sigma
__global__ void sn_finalize_kernel(const float* __restrict__ wv, float* __restrict__ u, float* __restrict__ out, int rows,
                                   int train, float eps) {
  pdl_sync();
  __shared__ double sh[32];
  __shared__ float s_inv;
  double acc = 0.0;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) acc += (double)wv[i] * wv[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sh[i];
    s_inv = 1.f / fmaxf((float)sqrt(s), eps);
  }
  __syncthreads();
  double dot = 0.0;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) {
    float ui = train ? wv[i] * s_inv : u[i];
    if (train) u[i] = ui;
    dot += (double)ui * wv[i];
  }
  dot = warp_sum(dot);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sh[i];
    out[0] = (float)s;
    out[1] = (float)(1.0 / s);
  }
}

// backward through sigma (u, v constants): dW = A - (<A, W> / sigma) * u v^T with A = (dL/dW_sn) / sigma already
// produced by the wgrad kernel (alpha = 1/sigma).
__global__ void sn_dot_kernel(const float* __restrict__ a, const float* __restrict__ w, long long n, double* __restrict__ out) {
  pdl_sync();
  double acc = 0.0;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x)
    acc += (double)a[i] * w[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}
__global__ void sn_rank1_kernel(float* __restrict__ a, const float* __restrict__ u, const float* __restrict__ v,
                                const double* __restrict__ dot, const float* __restrict__ sigma, int rows, int cols,
                                float* __restrict__ accum) {
  pdl_sync();
  const long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (long long)rows * cols) return;
  const float coef = (float)(dot[0]) * sigma[1];  // <A, W> / sigma
  const float r = a[i] - coef * u[i / cols] * v[i % cols];
  if (accum) accum[i] += r;  // straight into the optimizer's gradient bucket (a stays the per-pass scratch)
  else a[i] = r;
}

// ---- all spectrally-normalised layers of a network in one launch per phase (the layers are independent; only the four
// phases of one layer depend on each other): 4 launches per Discriminator pass instead of 20.
constexpr int kSnMaxLayers = 8;
struct SnBatch {
  const float* w[kSnMaxLayers];
  float* u[kSnMaxLayers];
  float* v[kSnMaxLayers];
  float* u_used[kSnMaxLayers];  // optional copies of the u / v this pass ends with (what its backward needs)
  float* v_used[kSnMaxLayers];
  float* t[kSnMaxLayers];
  float* wv[kSnMaxLayers];
  double* sumsq[kSnMaxLayers];
  float* out[kSnMaxLayers];
  int rows[kSnMaxLayers], cols[kSnMaxLayers];
  int train;
  float eps;
};
__global__ void sn_wt_u_batch_kernel(const SnBatch b, int rows_per_block) {
  pdl_sync();
  const int l = blockIdx.z, rows = b.rows[l], cols = b.cols[l];
  const int col = blockIdx.x * blockDim.x + threadIdx.x;
  const int r0 = blockIdx.y * rows_per_block;
  if (col >= cols || r0 >= rows) return;
  const int r1 = min(r0 + rows_per_block, rows);
  const float* __restrict__ w = b.w[l];
  const float* __restrict__ u = b.u[l];
  float acc = 0.f;
  for (int r = r0; r < r1; ++r) acc += __ldg(w + (long long)r * cols + col) * __ldg(u + r);
  atomicAdd(b.t[l] + col, acc);
}
__global__ void sn_sumsq_batch_kernel(const SnBatch b) {
  pdl_sync();
  const int l = blockIdx.y, n = b.cols[l];
  const float* __restrict__ t = b.t[l];
  double acc = 0.0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < n; i += gridDim.x * blockDim.x) acc += (double)t[i] * t[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0 && acc != 0.0) atomicAdd(b.sumsq[l], acc);
}
// one BLOCK per row (a warp per row walked d5's 6400 columns in 200 dependent trips: 100 us, latency-bound)
__global__ void __launch_bounds__(256) sn_w_v_batch_kernel(const SnBatch b) {
  pdl_sync();
  __shared__ float sh[8];
  const int l = blockIdx.y, rows = b.rows[l], cols = b.cols[l];
  const int row = blockIdx.x;
  if (row >= rows) return;
  float inv = 1.f;
  if (b.train) inv = 1.f / fmaxf((float)sqrt(*b.sumsq[l]), b.eps);
  const float* __restrict__ src = b.train ? b.t[l] : b.v[l];
  const float* __restrict__ w = b.w[l] + (long long)row * cols;
  float acc = 0.f;
  for (int c = threadIdx.x; c < cols; c += 256) {
    const float vv = src[c] * inv;
    acc += __ldg(w + c) * vv;
    if (row == 0) {  // row 0's block also publishes the new v (and the copy this pass's backward reads)
      if (b.train) b.v[l][c] = vv;
      if (b.v_used[l]) b.v_used[l][c] = vv;
    }
  }
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = sh[0];
#pragma unroll
    for (int i = 1; i < 8; ++i) t += sh[i];
    b.wv[l][row] = t;
  }
}
__global__ void sn_finalize_batch_kernel(const SnBatch b) {
  pdl_sync();
  __shared__ double sh[32];
  __shared__ float s_inv;
  const int l = blockIdx.x, rows = b.rows[l];
  const float* __restrict__ wv = b.wv[l];
  double acc = 0.0;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) acc += (double)wv[i] * wv[i];
  acc = warp_sum(acc);
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sh[i];
    s_inv = 1.f / fmaxf((float)sqrt(s), b.eps);
  }
  __syncthreads();
  double dot = 0.0;
  for (int i = threadIdx.x; i < rows; i += blockDim.x) {
    const float ui = b.train ? wv[i] * s_inv : b.u[l][i];
    if (b.train) b.u[l][i] = ui;
    if (b.u_used[l]) b.u_used[l][i] = ui;
    dot += (double)ui * wv[i];
  }
  dot = warp_sum(dot);
  __syncthreads();
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = dot;
  __syncthreads();
  if (threadIdx.x == 0) {
    double s = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) s += sh[i];
    b.out[l][0] = (float)s;
    b.out[l][1] = (float)(1.0 / s);
  }
}

}  // namespace uegan

using namespace uegan;

extern "C" int uegan_spectral_sigma_batch(int32_t count, const float* const* w, float* const* u, float* const* v,
                                          const int32_t* rows, const int32_t* cols, int32_t train, float* const* sigma_out,
                                          float* const* ws, float* const* u_used, float* const* v_used, void* stream) {
  UEGAN_CHECK(count >= 1 && count <= kSnMaxLayers && w && u && v && rows && cols && sigma_out && ws,
              "spectral_sigma_batch: 1..%d layers, non-null tables", kSnMaxLayers);
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  SnBatch b;
  memset(&b, 0, sizeof(b));
  int max_rows = 0, max_cols = 0;
  for (int l = 0; l < count; ++l) {
    UEGAN_CHECK(w[l] && u[l] && v[l] && sigma_out[l] && ws[l], "spectral_sigma_batch: null pointer (layer %d)", l);
    b.w[l] = w[l]; b.u[l] = u[l]; b.v[l] = v[l]; b.out[l] = sigma_out[l];
    b.u_used[l] = u_used ? u_used[l] : nullptr;
    b.v_used[l] = v_used ? v_used[l] : nullptr;
    b.rows[l] = rows[l]; b.cols[l] = cols[l];
    // ws layout as uegan_spectral_sigma: t[cols] | wv[rows] | sumsq (double, 8-byte aligned)
    b.t[l] = ws[l];
    b.wv[l] = ws[l] + cols[l];
    b.sumsq[l] = reinterpret_cast<double*>(ws[l] + ((cols[l] + rows[l] + 1) / 2) * 2);
    if (rows[l] > max_rows) max_rows = rows[l];
    if (cols[l] > max_cols) max_cols = cols[l];
    if (train) {
      UEGAN_CUDA(cudaMemsetAsync(b.t[l], 0, sizeof(float) * cols[l], st));
      UEGAN_CUDA(cudaMemsetAsync(b.sumsq[l], 0, sizeof(double), st));
    }
  }
  b.train = train;
  b.eps = 1e-12f;
  if (train) {
    const int rpb = 32;
    dim3 grid((max_cols + 127) / 128, (max_rows + rpb - 1) / rpb, count);
    launch_pdl(sn_wt_u_batch_kernel, grid, 128, 0, st, b, rpb);
    int sb = (max_cols + 1023) / 1024;
    if (sb > 32) sb = 32;
    launch_pdl(sn_sumsq_batch_kernel, dim3(sb, count), 256, 0, st, b);
  }
  launch_pdl(sn_w_v_batch_kernel, dim3(max_rows, count), 256, 0, st, b);
  launch_pdl(sn_finalize_batch_kernel, count, 256, 0, st, b);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int uegan_spectral_sigma(const float* w, float* u, float* v, int32_t rows, int32_t cols, int32_t train,
                                    float* sigma_out, float* ws, void* stream) {
  UEGAN_CHECK(w && u && v && sigma_out && ws, "spectral_sigma: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  // ws layout: t[cols] | wv[rows] | sumsq (double, 8-byte aligned)
  float* t = ws;
  float* wv = ws + cols;
  double* sumsq = reinterpret_cast<double*>(ws + ((cols + rows + 1) / 2) * 2);
  const float eps = 1e-12f;
  if (train) {
    UEGAN_CUDA(cudaMemsetAsync(t, 0, sizeof(float) * cols, st));
    UEGAN_CUDA(cudaMemsetAsync(sumsq, 0, sizeof(double), st));
    const int rpb = 32;
    dim3 grid((cols + 127) / 128, (rows + rpb - 1) / rpb);
    launch_pdl(sn_wt_u_kernel, grid, 128, 0, st, w, u, t, rows, cols, rpb);
    launch_pdl(sn_sumsq_kernel, (cols + 1023) / 1024 < 32 ? (cols + 1023) / 1024 : 32, 256, 0, st, t, cols, sumsq);
  }
  launch_pdl(sn_w_v_kernel, (rows * 32 + 255) / 256, 256, 0, st, w, t, sumsq, v, wv, rows, cols, train, eps);
  launch_pdl(sn_finalize_kernel, 1, 256, 0, st, wv, u, sigma_out, rows, train, eps);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

extern "C" int uegan_spectral_bwd(float* grad_inout, const float* w, const float* u, const float* v, const float* sigma,
                                  int32_t rows, int32_t cols, double* ws, float* accum_out, void* stream) {
  UEGAN_CHECK(grad_inout && w && u && v && sigma && ws, "spectral_bwd: null pointer");
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  const long long n = (long long)rows * cols;
  UEGAN_CUDA(cudaMemsetAsync(ws, 0, sizeof(double), st));
  int blocks = (int)((n + 1023) / 1024);
  if (blocks > 256) blocks = 256;
  launch_pdl(sn_dot_kernel, blocks, 256, 0, st, grad_inout, w, n, ws);
  launch_pdl(sn_rank1_kernel, (unsigned)((n + 255) / 256), 256, 0, st, grad_inout, u, v, ws, sigma, rows, cols, accum_out);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}
