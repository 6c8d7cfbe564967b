// Host-side helpers shared by the C-ABI translation units: error reporting, tensor geometry, TMA encode.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/uegan_sm100.h"

namespace uegan {

int set_error(const char* fmt, ...);  // returns -1
const char* get_error();

#define UEGAN_CHECK(cond, ...)               \
  do {                                       \
    if (!(cond)) return set_error(__VA_ARGS__); \
  } while (0)

#define UEGAN_CUDA(expr)                                                                      \
  do {                                                                                        \
    cudaError_t _e = (expr);                                                                  \
    if (_e != cudaSuccess) return set_error("%s failed: %s", #expr, cudaGetErrorString(_e));  \
  } while (0)

inline int dtype_size(int dtype) { return dtype == UEGAN_F32 ? 4 : 2; }
inline bool dtype_ok(int dtype) { return dtype == UEGAN_F32 || dtype == UEGAN_BF16 || dtype == UEGAN_F16; }
inline int64_t t_wp(const uegan_tensor& t) { return (int64_t)t.w + 2 * t.halo; }
inline int64_t t_hp(const uegan_tensor& t) { return (int64_t)t.h + 2 * t.halo; }
inline int64_t t_elems(const uegan_tensor& t) { return (int64_t)t.n * t_hp(t) * t_wp(t) * t.c; }

// cuTensorMapEncodeTiled through the runtime's driver entry point (no link-time libcuda dependency).
int encode_tiled(CUtensorMap* map, CUtensorMapDataType dt, uint32_t rank, void* base, const uint64_t* dims,
                 const uint64_t* strides_bytes /* rank-1 */, const uint32_t* box, CUtensorMapSwizzle swz);

int num_sms();

// Programmatic dependent launch (opt-in, UEGAN_PDL=1; measured slower for the two-stream training graph, see pdl_enabled):
// every kernel of this library starts with pdl_wait() (common.cuh:
// griddepcontrol.wait -- returns once the preceding kernel of the stream has completed and its writes are visible) before
// its first global-memory access and then releases its own successor (griddepcontrol.launch_dependents), so the next
// kernel's CTAs are placed and run their on-chip prologue (smem carve-up, mbarrier init, TMEM allocation) while this one
// drains, instead of after a full kernel boundary.  Correct by induction: a kernel's wait returns only after its
// predecessor completed, whose own wait returned only after ITS predecessor completed.  Works inside stream capture
// (programmatic graph edges).  Kernels launched without the attribute (ATen, the peer-memory optimiser) serialise fully.
bool pdl_enabled();
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// One host-mapped word the kernels' bounded waits write their code to before trapping.
unsigned int* error_sink_host();
unsigned int* error_sink_device();

}  // namespace uegan
