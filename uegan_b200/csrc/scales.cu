// Per-tensor power-of-two scales (uegan_tensor.scale) maintained on the device -- see include/uegan_sm100.h.
// One block per table entry: strided sample of the buffer (16-byte loads), block max of |stored value|, then
//   scale <- 2^floor(log2(target / (amax_stored / scale)))
// HBM-side cost: <= max_samples elements per tensor per pass (default 64 Ki): negligible next to the pass itself.
#include "common.cuh"
#include "host_util.h"

namespace uegan {

__global__ void __launch_bounds__(1024) scale_update_kernel(const uegan_scale_entry* __restrict__ tab, float target,
                                                           int max_samples) {
  pdl_sync();
  __shared__ float sh[32];
  __shared__ int sh_bad[32];
  const uegan_scale_entry e = tab[blockIdx.x];
  const int es = e.dtype == UEGAN_F32 ? 4 : 2;
  const long long nvec = (e.numel * es) / 16;  // whole 16-byte vectors (buffers are 16-byte aligned; tails are slack)
  long long want = (long long)max_samples * es / 16;
  if (want < 1) want = 1;
  const long long stride = nvec > want ? nvec / want : 1;
  float amax = 0.f;
  int bad = 0;
  // (1024 threads, four strided loads in flight per thread: the sample is a latency-bound gather -- one 256-thread block
  // walking it serially took 38 us per tensor, 0.6 ms per step)
  const uint4* base = reinterpret_cast<const uint4*>(e.data);
  const long long nsamp = (nvec + stride - 1) / stride;
  for (long long v0 = threadIdx.x; v0 < nsamp; v0 += 4ll * blockDim.x) {
    uint4 qq[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const long long v = v0 + (long long)u * blockDim.x;
      qq[u] = v < nsamp ? __ldg(base + v * stride) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
    const uint4 q = qq[u];
    const uint32_t w[4] = {q.x, q.y, q.z, q.w};
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      if (e.dtype == UEGAN_F32) {
        const float f = fabsf(__uint_as_float(w[i]));
        if (f <= 3.0e38f) amax = fmaxf(amax, f); else bad = 1;
      } else if (e.dtype == UEGAN_F16) {
        const __half2 h = *reinterpret_cast<const __half2*>(&w[i]);
        const float a = fabsf(__low2float(h)), b = fabsf(__high2float(h));
        // (stores saturate at 65504: a sample at the limit counts as overflow)
        if (a < 65504.f) amax = fmaxf(amax, a); else bad = 1;
        if (b < 65504.f) amax = fmaxf(amax, b); else bad = 1;
      } else {
        const __nv_bfloat162 h = *reinterpret_cast<const __nv_bfloat162*>(&w[i]);
        const float a = fabsf(__low2float(h)), b = fabsf(__high2float(h));
        if (a <= 3.0e38f) amax = fmaxf(amax, a); else bad = 1;
        if (b <= 3.0e38f) amax = fmaxf(amax, b); else bad = 1;
      }
    }
    }
  }
  for (int o = 16; o > 0; o >>= 1) {
    amax = fmaxf(amax, __shfl_xor_sync(0xffffffffu, amax, o));
    bad |= __shfl_xor_sync(0xffffffffu, bad, o);
  }
  if ((threadIdx.x & 31) == 0) { sh[threadIdx.x >> 5] = amax; sh_bad[threadIdx.x >> 5] = bad; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int i = 1; i < (int)(blockDim.x >> 5); ++i) { amax = fmaxf(amax, sh[i]); bad |= sh_bad[i]; }
    float s = *e.scale;
    if (!(s > 0.f) || !(s <= 3.0e38f)) s = 1.f;
    if (bad) {
      s *= 0.00390625f;  // an inf / nan was stored: back off by 2^8, the next pass measures again
    } else if (amax == 0.f) {
      // nothing but zeros in the sample: either the tensor IS zero (then its scale does not matter) or its values
      // underflowed fp16 (orthogonal(0.02) weights shrink activations by ~50x per layer): climb by 2^8 per pass
      if (e.dtype != UEGAN_F32 && s < 1.0e18f) s *= 256.f;
    } else {
      const float true_amax = e.reserved ? amax : amax / s;  // reserved != 0: the buffer holds TRUE values (fp32 master weights)
      int ex;
      frexpf(target / true_amax, &ex);       // target / true_amax = m * 2^ex, m in [0.5, 1)  ->  floor(log2) = ex - 1
      ex -= 1;
      if (ex > 100) ex = 100;
      if (ex < -100) ex = -100;
      s = ldexpf(1.f, ex);
    }
    *e.scale = s;
  }
}

}  // namespace uegan

using namespace uegan;

extern "C" int uegan_scale_update(const uegan_scale_entry* entries_dev, int32_t count, float target, int32_t max_samples,
                                  void* stream) {
  UEGAN_CHECK(entries_dev || count == 0, "scale_update: null table");
  UEGAN_CHECK(target > 0.f && max_samples > 0, "scale_update: bad target / sample count");
  if (count <= 0) return 0;
  launch_pdl(scale_update_kernel, (unsigned)count, 1024, 0, static_cast<cudaStream_t>(stream), entries_dev, target, max_samples);
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}
