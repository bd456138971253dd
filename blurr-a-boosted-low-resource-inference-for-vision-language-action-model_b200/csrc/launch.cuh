// Kernel launch helper: every kernel of the control step is launched with programmatic dependent
// launch (PDL) so that the launch latency and prologue of kernel N+1 overlap the execution of
// kernel N, both in eager streams and inside the captured CUDA graph.  Kernels call
// `pdl_wait()` before touching memory written by their predecessor and `pdl_trigger()` once the
// successor may start launching.
#pragma once

#include <cuda_runtime.h>

namespace blurr {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;\n" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;\n" ::: "memory"); }

bool pdl_enabled();
void pdl_set_enabled(bool on);

template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream,
                                 Args... args) {
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = stream;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

}  // namespace blurr
