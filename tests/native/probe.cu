// Hardware probe (test-only entry point): how does tcgen05.mma address a K-major SWIZZLE_128B operand whose start is
// NOT the 1024-byte-aligned origin of the swizzle pattern?  The answer decides whether a convolution can stage one
// halo'd patch of the input in shared memory and feed every filter tap as a shifted descriptor window (DESIGN.md,
// "smem window reuse") instead of re-loading the tile once per tap.
//
// A: a_rows x 32 fp32, row r lives at smem_base + r*128 with its 16-byte chunks XOR-swizzled by ((addr >> 7) & 7)
//    (exactly what a SWIZZLE_128B TMA box with a 1024-aligned destination produces).
// B: n x 32 fp32, canonical K-major SWIZZLE_128B at a 1024-aligned address.
// D[m][j] = sum_k A[row(m)][k] * B[j][k],  row(m) = row_shift + (m / 8) * (sbo_bytes / 128) + m % 8  -- if the
// hardware applies the swizzle to absolute address bits (or honours base_offset).
#include "../../uegan_b200/csrc/common.cuh"
#include "../../uegan_b200/csrc/host_util.h"

namespace uegan {

__global__ void __launch_bounds__(128, 1)
probe_umma_window_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ out, int a_rows,
                         int n, int row_shift, int base_offset, int sbo_bytes, unsigned int* err_sink) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  __shared__ __align__(8) uint64_t done_bar;
  __shared__ uint32_t tmem_base_smem;
  uint8_t* sa = smem;
  uint8_t* sb = smem + ((a_rows * 128 + 1023) / 1024) * 1024;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // fill A and B with the software swizzle
  for (int i = threadIdx.x; i < a_rows * 8; i += blockDim.x) {
    const int r = i / 8, ch = i % 8;
    const uint32_t addr = smem_u32(sa) + r * 128;
    const int pch = ch ^ ((addr >> 7) & 7);
    *reinterpret_cast<float4*>(sa + r * 128 + pch * 16) = *reinterpret_cast<const float4*>(a + r * 32 + ch * 4);
  }
  for (int i = threadIdx.x; i < n * 8; i += blockDim.x) {
    const int r = i / 8, ch = i % 8;
    const int pch = ch ^ (r & 7);
    *reinterpret_cast<float4*>(sb + r * 128 + pch * 16) = *reinterpret_cast<const float4*>(b + r * 32 + ch * 4);
  }
  fence_proxy_async();
  if (warp == 0 && elect_one()) {
    mbar_init(&done_bar, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc(&tmem_base_smem, 256);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = tmem_base_smem;
  if (warp == 0 && elect_one()) {
    const uint32_t idesc = make_instr_desc(UMMA_TF32, 128, n);
    const uint32_t a_addr = smem_u32(sa) + row_shift * 128;
    const uint32_t b_addr = smem_u32(sb);
    for (int k = 0; k < 4; ++k) {
      const uint64_t da = make_smem_desc(a_addr + k * 32, 16, sbo_bytes, UMMA_LAYOUT_SW128, base_offset);
      const uint64_t db = make_smem_desc(b_addr + k * 32, 16, 1024, UMMA_LAYOUT_SW128);
      umma_ss<1>(tmem_base, da, db, idesc, k != 0);
    }
    umma_commit(&done_bar);
  }
  __syncwarp();
  mbar_wait(&done_bar, 0, 0x900, err_sink);
  tcgen05_fence_after();
  const int m = warp * 32 + lane;
  for (int c0 = 0; c0 < n; c0 += 16) {
    uint32_t r[16];
    tmem_ld16(tmem_base + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int i = 0; i < 16; ++i) out[m * n + c0 + i] = __uint_as_float(r[i]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, 256);
}

}  // namespace uegan

using namespace uegan;

extern "C" int uegan_probe_umma_window(const float* a, const float* b, float* out, int32_t a_rows, int32_t n,
                                       int32_t kchunks, int32_t row_shift, int32_t base_offset, int32_t sbo_bytes,
                                       void* stream) {
  UEGAN_CHECK(a && b && out, "probe: null pointer");
  UEGAN_CHECK(kchunks == 1, "probe: kchunks must be 1");
  UEGAN_CHECK(n % 16 == 0 && n >= 16 && n <= 256, "probe: bad n");
  UEGAN_CHECK(sbo_bytes % 128 == 0, "probe: sbo must be a multiple of 128");
  const int need_rows = row_shift + 15 * (sbo_bytes / 128) + 8;
  UEGAN_CHECK(a_rows >= need_rows, "probe: a_rows %d < %d", a_rows, need_rows);
  const int smem = ((a_rows * 128 + 1023) / 1024) * 1024 + n * 128 + 2048;
  UEGAN_CHECK(smem <= 210 * 1024, "probe: too much smem");
  static bool attr_set = false;
  if (!attr_set) {
    UEGAN_CUDA(cudaFuncSetAttribute(probe_umma_window_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 210 * 1024));
    attr_set = true;
  }
  probe_umma_window_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(a, b, out, a_rows, n, row_shift,
                                                                               base_offset, sbo_bytes,
                                                                               error_sink_device());
  UEGAN_CUDA(cudaGetLastError());
  return 0;
}
