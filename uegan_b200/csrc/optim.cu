// Adam over ONE flat fp32 bucket per network (torch.optim.Adam semantics, trainer.py:337-338: coupled weight decay 1e-4,
// eps 1e-8, no amsgrad), fused with the data-parallel gradient reduction:
//   * uegan_adam_step        g = this rank's flat gradient bucket (already reduced, or single GPU)
//   * uegan_adam_step_peers  g = sum over the W ranks' buckets read DIRECTLY from peer memory over NVLink / NVSwitch
//                            (symmetric-memory pointers), summed in rank order -- every rank computes bit-identical
//                            updates, no NCCL call, no reduced copy of the gradients is ever materialised
// The step counter and the learning rate live in device memory (CUDA-graph replays advance them without host work).
// HBM-bound: 4 streams read (p, g x W, m, v), 3 written, 16 bytes per thread per access.
#include "common.cuh"
#include "host_util.h"

namespace uegan {

// state[0] = step (float, as torch's capturable Adam keeps it), state[1] = lr / (1 - beta1^step), state[2] = 1 / sqrt(1 - beta2^step)
__global__ void adam_tick_kernel(float* __restrict__ state, const float* __restrict__ lr, float beta1, float beta2) {
  const float step = state[0] + 1.f;
  state[0] = step;
  const double bc1 = 1.0 - pow((double)beta1, (double)step);
  const double bc2 = 1.0 - pow((double)beta2, (double)step);
  state[1] = (float)((double)lr[0] / bc1);
  state[2] = (float)(1.0 / sqrt(bc2));
}

constexpr int kMaxPeers = 16;
struct PeerPtrs {
  const float* g[kMaxPeers];
};

template <int kPeers>  // 0: single local bucket in ptrs.g[0]
__global__ void __launch_bounds__(256) adam_flat_kernel(float* __restrict__ p, PeerPtrs ptrs, int world, float* __restrict__ m,
                                                       float* __restrict__ v, long long n4, long long n,
                                                       const float* __restrict__ state, float beta1, float beta2, float eps,
                                                       float wd) {
  const float step_size = state[1], inv_sqrt_bc2 = state[2];
  const long long stride = (long long)gridDim.x * blockDim.x;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const long long o = i * 4;
    float pv[4], gv[4], mv[4], vv[4];
    if (o + 3 < n) {
      const float4 a = *reinterpret_cast<const float4*>(p + o);
      pv[0] = a.x; pv[1] = a.y; pv[2] = a.z; pv[3] = a.w;
      const float4 b = *reinterpret_cast<const float4*>(m + o);
      mv[0] = b.x; mv[1] = b.y; mv[2] = b.z; mv[3] = b.w;
      const float4 c = *reinterpret_cast<const float4*>(v + o);
      vv[0] = c.x; vv[1] = c.y; vv[2] = c.z; vv[3] = c.w;
      const float4 g0 = *reinterpret_cast<const float4*>(ptrs.g[0] + o);
      gv[0] = g0.x; gv[1] = g0.y; gv[2] = g0.z; gv[3] = g0.w;
      if (kPeers) {
        for (int r = 1; r < world; ++r) {  // fixed rank order: identical sums on every rank
          const float4 gr = *reinterpret_cast<const float4*>(ptrs.g[r] + o);
          gv[0] += gr.x; gv[1] += gr.y; gv[2] += gr.z; gv[3] += gr.w;
        }
      }
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        const bool ok = o + k < n;
        pv[k] = ok ? p[o + k] : 0.f; mv[k] = ok ? m[o + k] : 0.f; vv[k] = ok ? v[o + k] : 0.f;
        float g = ok ? ptrs.g[0][o + k] : 0.f;
        if (kPeers) for (int r = 1; r < world; ++r) g += ok ? ptrs.g[r][o + k] : 0.f;
        gv[k] = g;
      }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const float g = fmaf(wd, pv[k], gv[k]);                 // grad + weight_decay * param
      mv[k] = fmaf(1.f - beta1, g - mv[k], mv[k]);            // exp_avg.lerp_(grad, 1 - beta1)
      vv[k] = fmaf(1.f - beta2, g * g, beta2 * vv[k]);        // exp_avg_sq.mul_(beta2).addcmul_(grad, grad, 1 - beta2)
      const float denom = fmaf(sqrtf(vv[k]), inv_sqrt_bc2, eps);
      pv[k] = pv[k] - step_size * (mv[k] / denom);
    }
    if (o + 3 < n) {
      *reinterpret_cast<float4*>(p + o) = make_float4(pv[0], pv[1], pv[2], pv[3]);
      *reinterpret_cast<float4*>(m + o) = make_float4(mv[0], mv[1], mv[2], mv[3]);
      *reinterpret_cast<float4*>(v + o) = make_float4(vv[0], vv[1], vv[2], vv[3]);
    } else {
#pragma unroll
      for (int k = 0; k < 4; ++k)
        if (o + k < n) { p[o + k] = pv[k]; m[o + k] = mv[k]; v[o + k] = vv[k]; }
    }
  }
}

// out[i] = sum over ranks r (in rank order) of peers[r][i], i < count <= 64: the batch-global partial sums of the
// relativistic GAN loss (losses.py:351-360 means over the WHOLE batch) without a NCCL call.
struct PeerPtrsF64 {
  const double* p[kMaxPeers];
};
__global__ void peer_sum_f64_kernel(double* __restrict__ out, PeerPtrsF64 ptrs, int world, int count) {
  const int i = threadIdx.x;
  if (i >= count) return;
  double acc = 0.0;
  for (int r = 0; r < world; ++r) acc += ptrs.p[r][i];
  out[i] = acc;
}

static int adam_launch(float* p, const PeerPtrs& ptrs, int world, float* m, float* v, long long n, float* state,
                       const float* lr, float beta1, float beta2, float eps, float wd, cudaStream_t st) {
  const long long n4 = (n + 3) / 4;
  long long blocks = (n4 + 255) / 256;
  const long long cap = 8LL * num_sms();
  if (blocks > cap) blocks = cap;
  adam_tick_kernel<<<1, 1, 0, st>>>(state, lr, beta1, beta2);
  if (world > 1)
    adam_flat_kernel<1><<<(unsigned)blocks, 256, 0, st>>>(p, ptrs, world, m, v, n4, n, state, beta1, beta2, eps, wd);
  else
    adam_flat_kernel<0><<<(unsigned)blocks, 256, 0, st>>>(p, ptrs, 1, m, v, n4, n, state, beta1, beta2, eps, wd);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace uegan

using namespace uegan;

extern "C" {

int uegan_adam_step(float* param, const float* grad, float* exp_avg, float* exp_avg_sq, int64_t n, float* state3,
                    const float* lr_dev, float beta1, float beta2, float eps, float weight_decay, void* stream) {
  UEGAN_CHECK(param && grad && exp_avg && exp_avg_sq && state3 && lr_dev && n > 0, "adam_step: null pointer");
  UEGAN_CHECK(((uintptr_t)param | (uintptr_t)grad | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0,
              "adam_step: buckets must be 16-byte aligned");
  PeerPtrs ptrs;
  memset(&ptrs, 0, sizeof(ptrs));
  ptrs.g[0] = grad;
  return adam_launch(param, ptrs, 1, exp_avg, exp_avg_sq, n, state3, lr_dev, beta1, beta2, eps, weight_decay,
                     static_cast<cudaStream_t>(stream));
}

int uegan_adam_step_peers(float* param, const float* const* peer_grads_host, int32_t world, float* exp_avg,
                          float* exp_avg_sq, int64_t n, float* state3, const float* lr_dev, float beta1, float beta2,
                          float eps, float weight_decay, void* stream) {
  UEGAN_CHECK(param && peer_grads_host && exp_avg && exp_avg_sq && state3 && lr_dev && n > 0, "adam_step_peers: null pointer");
  UEGAN_CHECK(world >= 1 && world <= kMaxPeers, "adam_step_peers: world %d unsupported (max %d)", world, kMaxPeers);
  PeerPtrs ptrs;
  memset(&ptrs, 0, sizeof(ptrs));
  for (int r = 0; r < world; ++r) {
    UEGAN_CHECK(peer_grads_host[r] && (uintptr_t)peer_grads_host[r] % 16 == 0, "adam_step_peers: bad peer pointer %d", r);
    ptrs.g[r] = peer_grads_host[r];
  }
  UEGAN_CHECK(((uintptr_t)param | (uintptr_t)exp_avg | (uintptr_t)exp_avg_sq) % 16 == 0,
              "adam_step_peers: buckets must be 16-byte aligned");
  return adam_launch(param, ptrs, world, exp_avg, exp_avg_sq, n, state3, lr_dev, beta1, beta2, eps, weight_decay,
                     static_cast<cudaStream_t>(stream));
}

int uegan_memset_zero(void* ptr, size_t bytes, void* stream) {
  UEGAN_CHECK(ptr || bytes == 0, "memset_zero: null pointer");
  if (bytes) UEGAN_CUDA(cudaMemsetAsync(ptr, 0, bytes, static_cast<cudaStream_t>(stream)));
  return 0;
}

int uegan_capture_status(void* stream) {
  cudaStreamCaptureStatus st = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing(static_cast<cudaStream_t>(stream), &st) != cudaSuccess) return -1;
  return (int)st;  // 0 none, 1 active, 2 invalidated
}

int uegan_peer_sum_f64(double* out, const double* const* peers_host, int32_t world, int32_t count, void* stream) {
  UEGAN_CHECK(out && peers_host && world >= 1 && world <= kMaxPeers && count >= 1 && count <= 64,
              "peer_sum_f64: bad arguments (world %d, count %d)", world, count);
  PeerPtrsF64 ptrs;
  memset(&ptrs, 0, sizeof(ptrs));
  for (int r = 0; r < world; ++r) {
    UEGAN_CHECK(peers_host[r], "peer_sum_f64: null peer pointer %d", r);
    ptrs.p[r] = peers_host[r];
  }
  peer_sum_f64_kernel<<<1, 64, 0, static_cast<cudaStream_t>(stream)>>>(out, ptrs, world, count);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}

}  // extern "C"
