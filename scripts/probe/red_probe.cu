// Probe (B200): throughput of global fp32 reductions in the access pattern of the warp
// backward's grad_input scatter -- scalar RED.F32, RED.v4.F32, and TMA bulk reduce-add from
// shared memory.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o red_probe red_probe.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

constexpr int H = 1088, W = 1920, C = 64;

// pattern 1: thread = pixel, 4 scalar REDs per channel (NW, NE, SW, SE with a 1-px shift)
__global__ void k_scalar(float* g, int nch) {
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W - 2 || y >= H - 2) return;
    float* p = g + (size_t)y * W + x;
    for (int c = 0; c < nch; ++c) {
        atomicAdd(p, 1.0f);
        atomicAdd(p + 1, 1.0f);
        atomicAdd(p + W, 1.0f);
        atomicAdd(p + W + 1, 1.0f);
        p += (size_t)H * W;
    }
}
__device__ __forceinline__ void red4(float* p, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(a), "f"(b), "f"(c), "f"(d) : "memory");
}
// pattern 2: thread = 4 consecutive pixels; 3 rows x 2 aligned quads of RED.v4 per channel
__global__ void k_v4(float* g, int nch, int nquads, int nrows) {
    const int x = (blockIdx.x * 32 + threadIdx.x) * 4, y = blockIdx.y * 8 + threadIdx.y;
    if (x >= W - 8 || y >= H - 3) return;
    float* p = g + (size_t)y * W + x;
    for (int c = 0; c < nch; ++c) {
        for (int r = 0; r < nrows; ++r)
            for (int q = 0; q < nquads; ++q) red4(p + r * W + 4 * q, 1.f, 1.f, 1.f, 1.f);
        p += (size_t)H * W;
    }
}
// pattern 3: CTA = 64x16 tile; per channel a [24 rows][96] shared-memory box is added to global
// with one bulk reduce per row (384 B), box origin = tile origin (boxes of neighbours overlap)
__global__ void __launch_bounds__(256) k_bulk(float* g, int nch, int rows, int bw) {
    extern __shared__ __align__(128) float sm[];
    for (int i = threadIdx.x; i < rows * bw; i += 256) sm[i] = 1.0f;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int x0 = blockIdx.x * 64, y0 = blockIdx.y * 16;
    if (threadIdx.x < rows) {
        const int r = threadIdx.x;
        if (y0 + r < H && x0 + bw <= W) {
            float* p = g + (size_t)(y0 + r) * W + x0;
            const uint32_t s = (uint32_t)__cvta_generic_to_shared(sm + r * bw);
            for (int c = 0; c < nch; ++c) {
                asm volatile("cp.reduce.async.bulk.global.shared::cta.bulk_group.add.f32 [%0], [%1], %2;"
                             ::"l"(p), "r"(s), "r"(bw * 4) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
                p += (size_t)H * W;
                if ((c & 7) == 7) asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
            }
            asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
        }
    }
}

int main() {
    float* g;
    const size_t n = (size_t)C * H * W;
    CK(cudaMalloc(&g, n * 4));
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    auto run = [&](const char* name, auto launch, double elems) {
        float best = 1e9;
        for (int it = 0; it < 5; ++it) {
            cudaMemsetAsync(g, 0, n * 4);
            cudaEventRecord(e0);
            launch();
            cudaEventRecord(e1);
            cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1);
            if (ms < best) best = ms;
        }
        cudaError_t e = cudaGetLastError();
        printf("%-44s %8.1f us  %7.2f G elem-adds/s  (%s)\n", name, best * 1e3, elems / best * 1e-6, cudaGetErrorString(e));
    };
    {
        float ms = 0; cudaEventRecord(e0); cudaMemsetAsync(g, 0, n * 4); cudaEventRecord(e1); cudaEventSynchronize(e1);
        cudaEventElapsedTime(&ms, e0, e1); printf("memset %zu MB: %.1f us\n", n * 4 >> 20, ms * 1e3);
    }
    dim3 b(32, 8);
    run("scalar RED x4 / px-ch", [&] { k_scalar<<<dim3(W / 32, H / 8), b>>>(g, C); }, 4.0 * H * W * C);
    run("RED.v4 3 rows x 2 quads / 4px-ch", [&] { k_v4<<<dim3(W / 128, H / 8), b>>>(g, C, 2, 3); }, 24.0 * H * (W / 4) * C);
    run("RED.v4 2 rows x 2 quads / 4px-ch", [&] { k_v4<<<dim3(W / 128, H / 8), b>>>(g, C, 2, 2); }, 16.0 * H * (W / 4) * C);
    run("RED.v4 1 row x 1 quad / 4px-ch", [&] { k_v4<<<dim3(W / 128, H / 8), b>>>(g, C, 1, 1); }, 4.0 * H * (W / 4) * C);
    cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    run("bulk reduce 24 rows x 96 / tile-ch", [&] { k_bulk<<<dim3(W / 64, H / 16), 256, 24 * 96 * 4>>>(g, C, 24, 96); }, 24.0 * 96 * (W / 64) * (H / 16) * C);
    run("bulk reduce 24 rows x 80 / tile-ch", [&] { k_bulk<<<dim3(W / 64, H / 16), 256, 24 * 80 * 4>>>(g, C, 24, 80); }, 24.0 * 80 * (W / 64) * (H / 16) * C);
    run("bulk reduce 16 rows x 64 / tile-ch (exact)", [&] { k_bulk<<<dim3(W / 64, H / 16), 256, 16 * 64 * 4>>>(g, C, 16, 64); }, 16.0 * 64 * (W / 64) * (H / 16) * C);
    // check: sum of one plane for the last pattern
    CK(cudaDeviceSynchronize());
    float h[8]; CK(cudaMemcpy(h, g + 5 * W + 64, 32, cudaMemcpyDeviceToHost));
    printf("sample %g %g\n", h[0], h[1]);
    return 0;
}
