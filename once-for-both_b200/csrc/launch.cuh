// launch.cuh — one launcher for every kernel of the library: programmatic dependent launch (PDL) + optional CTA clusters.
//
// The search step is ~255 kernels of 10-200 us each, replayed from a CUDA graph. Between two kernels of a stream the GPU
// otherwise idles for the launch latency plus the next kernel's prologue (mbarrier init, TMEM allocation, tensor-map
// prefetch): measured 0.7 ms per step (sum of kernel durations 13.2 ms vs 13.9 ms per step). With the programmatic stream
// serialisation attribute the NEXT kernel may be launched as soon as every CTA of the current one has executed
// griddepcontrol.launch_dependents (pdl_trigger(), first instruction of the persistent one-CTA-per-SM kernels): its CTAs
// become resident SM by SM while the current kernel drains, run their prologue, and block in griddepcontrol.wait (pdl_wait())
// until the current grid has completed and its memory is visible. Every kernel of the library calls pdl_wait() before it
// touches global memory; kernels that do not trigger explicitly trigger implicitly when they finish. Stream capture turns the
// attribute into programmatic edges of the CUDA graph.
// Measured on B200 (DeiT-S step, graph replay, same process): 14.14 -> 14.05 ms per step (-0.6 %): a resident CTA needs
// >200 KB of shared memory, so the next kernel's CTAs can only move in SM by SM as the current kernel drains - what overlaps is
// launch latency, not the prologue. Because the CUPTI durations of overlapped kernels include their wait (the per-family
// roofline of bench.py would read low), the attribute is OPT-IN: OFB_PDL=1. All kernels carry pdl_wait() either way (it is a
// no-op without a programmatic dependency).
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>

namespace ofb {

__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

inline bool pdl_enabled() {
    static int on = -1;
    if (on < 0) {
        const char* e = getenv("OFB_PDL");
        on = (e != nullptr && atoi(e) != 0) ? 1 : 0;
    }
    return on == 1;
}

template <typename... P, typename... A>
inline cudaError_t launch_k(void (*kernel)(P...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, int cluster, A... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid; cfg.blockDim = block; cfg.dynamicSmemBytes = smem; cfg.stream = stream;
    cudaLaunchAttribute attr[2];
    unsigned n = 0;
    if (pdl_enabled()) {
        attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[n].val.programmaticStreamSerializationAllowed = 1;
        ++n;
    }
    if (cluster > 1) {
        attr[n].id = cudaLaunchAttributeClusterDimension;
        attr[n].val.clusterDim.x = unsigned(cluster); attr[n].val.clusterDim.y = 1; attr[n].val.clusterDim.z = 1;
        ++n;
    }
    cfg.attrs = attr; cfg.numAttrs = n;
    return cudaLaunchKernelEx(&cfg, kernel, P(args)...);
}

}  // namespace ofb

#define OFB_LAUNCH(kernel, grid, block, smem, stream, ...) \
    (void)ofb::launch_k(kernel, dim3(grid), dim3(block), size_t(smem), stream, 1, __VA_ARGS__)
