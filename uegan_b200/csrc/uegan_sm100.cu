// Unity translation unit of libuegan_sm100.so (one nvcc invocation, no relocatable device code).
#include "host_util.cu"
#include "conv_fprop.cu"
#include "elementwise.cu"
#include "spectral.cu"
#include "losses.cu"
#include "backward.cu"
#include "conv_wgrad.cu"
#include "probe.cu"
