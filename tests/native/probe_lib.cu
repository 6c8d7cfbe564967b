// Unity translation unit of tests/native/libuegan_probe.so: hardware probes used by the test-suite only (not part of
// the product library).  Shares the product's device helpers (common.cuh) and host utilities by inclusion.
#include "../../uegan_b200/csrc/host_util.cu"
#include "probe.cu"

extern "C" int uegan_probe_device_error(void) {
  cudaError_t e = cudaDeviceSynchronize();
  unsigned int* sink = uegan::error_sink_host();
  const unsigned int v = sink ? *sink : 0;
  if (sink) *sink = 0;
  if (e != cudaSuccess) return v ? (int)v : -1;
  return (int)v;
}
