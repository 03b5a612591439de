// Programmatic dependent launch (PDL) for the latency-bound sampler step.
//
// A small call is ~120 kernels of 3-40 us per diffusion step; each boundary costs launch latency + block scheduling +
// the kernel's own prologue (mbarrier init, TMEM allocation, tensor-map fetch).  With PDL the next kernel's blocks are
// scheduled while the current kernel still runs (on SMs it leaves idle: most launches of a small call fill a fraction of the
// GPU) and wait at `griddepcontrol.wait` until the predecessor grid has completed and its writes are visible.  Every kernel
// of the step calls pdl_wait_and_trigger() before its first global read; the trigger comes after the wait, so exactly one
// kernel runs ahead.  Launched without the attribute (throughput mode, vocoder, fine-tune) both instructions are no-ops.
#pragma once
#include <utility>
#include <cuda_runtime.h>

namespace usb {

// engine.cu: thread-local switch, on only around the launches of a latency-mode sampler step
bool pdl_enabled();
void set_pdl(bool on);

__device__ __forceinline__ void pdl_wait_and_trigger() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <typename... KArgs, typename... Args>
inline cudaError_t launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

// same with a thread-block cluster of `cluster_x` CTAs along x (grid.x must be a multiple of it)
template <typename... KArgs, typename... Args>
inline cudaError_t launch_k_cluster(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, int cluster_x,
                                    Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = s;
    cudaLaunchAttribute at[2];
    at[0].id = cudaLaunchAttributeClusterDimension;
    at[0].val.clusterDim.x = cluster_x;
    at[0].val.clusterDim.y = 1;
    at[0].val.clusterDim.z = 1;
    at[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[1].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = pdl_enabled() ? 2 : 1;
    return cudaLaunchKernelEx(&cfg, kernel, std::forward<Args>(args)...);
}

}  // namespace usb
