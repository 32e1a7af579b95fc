// Kernel launch helper: every kernel of the forward goes through launch_k(), which (when enabled) marks the launch
// as a programmatic dependent launch.  All kernels start with pdl_sync(): they allow THEIR successor to be
// scheduled at once and then wait until their predecessor grid has completed and flushed - so data dependencies
// are exactly those of a serial stream, but the ~1-2 us launch / scheduling gap between two dependent kernels (120
// launches per forward, many of them 3-10 us long) is hidden.  Without the launch attribute the two PTX
// instructions are no-ops.  Works inside stream capture (CUDA >= 12.3 records programmatic edges).
#pragma once
#include <cuda_runtime.h>

#include <utility>

namespace lwb {

int& pdl_enabled();   // defined in tma_util.cu; set through lwdetr_set_option(engine, "pdl", v)

// Per-DEVICE one-time kernel attributes (tma_util.cu).  cudaFuncSetAttribute acts on the current device only, so the
// "done" record is keyed by (kernel, device): a second GPU used from the same process gets its own opt-in.
int ensure_max_dyn_smem(const void* kernel, int bytes);        // cudaFuncAttributeMaxDynamicSharedMemorySize
int current_device_sms();                                      // SM count of the current device (cached per device)

#ifdef __CUDACC__
__device__ __forceinline__ void pdl_sync() {
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  asm volatile("griddepcontrol.wait;" ::: "memory");
}
#endif

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl_enabled() ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, KArgs(std::forward<Args>(args))...);
}

}  // namespace lwb
