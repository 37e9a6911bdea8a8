// Shared device helpers for the deepsvc_b200 kernels (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#ifndef DSVC_NUM_SMS
#define DSVC_NUM_SMS 148  // B200: 2 dies x 74 SMs
#endif

#define DSVC_CHECK_ARG(cond) \
    do {                     \
        if (!(cond)) return (int)cudaErrorInvalidValue; \
    } while (0)

#define DSVC_RETURN_LAST() return (int)cudaGetLastError()

namespace dsvc {

// Programmatic dependent launch (PDL): the forward-path kernels are launched with the
// programmatic-stream-serialization attribute and begin with pdl_prologue().  A kernel's CTAs may
// then be scheduled while its predecessor in the stream (or captured graph) is still draining --
// they block in griddepcontrol.wait until the predecessor has COMPLETED and its writes are visible,
// so the stream's serial semantics are unchanged; what overlaps is the launch latency and the
// ramp of the ~20 few-microsecond launches of a frame.  launch_dependents right after the wait lets
// the successor be scheduled as soon as every CTA of this grid has started.
__device__ __forceinline__ void pdl_prologue() {
    asm volatile("griddepcontrol.wait;" ::: "memory");
    asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}

template <class... KArgs, class... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                              Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}

// The staged warp kernels configure their SMs for the maximum shared-memory carve-out.  A kernel
// that prefers another L1 / shared-memory split cannot become resident on such an SM until it
// drains, so the path's short kernels ask for the same carve-out (once per kernel and device)
// and can run next to the persistent feature warp instead of queueing behind it.
template <class K>
inline void prefer_max_shared_carveout(K kernel) {
    static unsigned long long done = 0;  // one flag per device (per instantiation)
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64 || ((done >> dev) & 1ull)) return;
    cudaFuncSetAttribute(kernel, cudaFuncAttributePreferredSharedMemoryCarveout, cudaSharedmemCarveoutMaxShared);
    done |= 1ull << dev;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum in fixed order (warp shuffle tree, then warp 0 over the per-warp
// partials). Result valid in thread 0. `smem` must hold blockDim.x/32 doubles.
__device__ __forceinline__ double block_sum(double v, double* smem) {
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
    v = warp_sum(v);
    if (lane == 0) smem[wid] = v;
    __syncthreads();
    const int nw = (blockDim.x + 31) >> 5;
    double r = 0.0;
    if (wid == 0) {
        r = lane < nw ? smem[lane] : 0.0;
        r = warp_sum(r);
    }
    return r;
}

__host__ __device__ __forceinline__ bool aligned16(const void* p) {
    return (reinterpret_cast<uintptr_t>(p) & 15u) == 0;
}

// streaming 128-bit load / store (no L1 allocation: every element is touched once)
__device__ __forceinline__ float4 ld_stream4(const float* p) {
    float4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.f32 {%0,%1,%2,%3}, [%4];"
                 : "=f"(r.x), "=f"(r.y), "=f"(r.z), "=f"(r.w)
                 : "l"(p));
    return r;
}
__device__ __forceinline__ void st_stream4(float* p, float4 v) {
    asm volatile("st.global.L1::no_allocate.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x),
                 "f"(v.y), "f"(v.z), "f"(v.w)
                 : "memory");
}
__device__ __forceinline__ void st_stream1(float* p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0], %1;" ::"l"(p), "f"(v) : "memory");
}

template <int IMM>
__device__ __forceinline__ void st_stream1_imm(float* p, float v) {
    asm volatile("st.global.L1::no_allocate.f32 [%0+%2], %1;" ::"l"(p), "f"(v), "n"(IMM) : "memory");
}

template <int IMM>
__device__ __forceinline__ void st_hint_imm(float* p, float v, uint64_t pol) {
    asm volatile("st.global.L1::no_allocate.L2::cache_hint.f32 [%0+%2], %1, %3;" ::"l"(p), "f"(v), "n"(IMM), "l"(pol) : "memory");
}
}  // namespace dsvc
