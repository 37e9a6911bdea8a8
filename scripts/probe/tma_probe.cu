// Minimal TMA 3-D tiled load probe (debug helper, not part of the product library).
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cudaTypedefs.h>
#include <cuda_runtime.h>
#include <vector>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int BW, int BH, int CC>
__global__ void probe(const __grid_constant__ CUtensorMap tmap, int x, int y, int z, float* out) {
    extern __shared__ __align__(128) unsigned char smem[];
    __shared__ __align__(8) uint64_t bar;
    float* buf = (float*)smem;
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(&bar)), "r"(1));
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(&bar)), "r"(BW * BH * CC * 4) : "memory");
        asm volatile("cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3, %4}], [%5];"
                     ::"r"(smem_u32(buf)), "l"((uint64_t)&tmap), "r"(x), "r"(y), "r"(z), "r"(smem_u32(&bar)) : "memory");
    }
    uint32_t done = 0;
    long long t0 = clock64();
    while (!done) {
        asm volatile("{\n.reg .pred p;\nmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\nselp.u32 %0, 1, 0, p;\n}\n" : "=r"(done) : "r"(smem_u32(&bar)), "r"(0) : "memory");
        if (clock64() - t0 > 2000000000ll) { if (threadIdx.x == 0) printf("timeout\n"); return; }
    }
    for (int i = threadIdx.x; i < BW * BH * CC; i += blockDim.x) out[i] = buf[i];
}

template <int BW, int BH, int CC>
int run(int W, int H, int P, int x, int y, int z) {
    std::vector<float> h((size_t)W * H * P);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (float)i;
    float *d, *o;
    cudaMalloc(&d, h.size() * 4);
    cudaMalloc(&o, BW * BH * CC * 4);
    cudaMemcpy(d, h.data(), h.size() * 4, cudaMemcpyHostToDevice);
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult q;
    cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &q);
    auto encode = (PFN_cuTensorMapEncodeTiled_v12000)fn;
    CUtensorMap tm;
    cuuint64_t gdim[3] = {(cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)P};
    cuuint64_t gstr[2] = {(cuuint64_t)W * 4, (cuuint64_t)W * H * 4};
    cuuint32_t box[3] = {BW, BH, CC};
    cuuint32_t es[3] = {1, 1, 1};
    CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, d, gdim, gstr, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                        CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    printf("box %dx%dx%d coord (%d,%d,%d): encode=%d ", BW, BH, CC, x, y, z, (int)r);
    cudaFuncSetAttribute(probe<BW, BH, CC>, cudaFuncAttributeMaxDynamicSharedMemorySize, BW * BH * CC * 4);
    probe<BW, BH, CC><<<1, 128, BW * BH * CC * 4>>>(tm, x, y, z, o);
    cudaError_t e = cudaDeviceSynchronize();
    printf("sync=%s ", cudaGetErrorString(e));
    if (e == cudaSuccess) {
        std::vector<float> res(BW * BH * CC);
        cudaMemcpy(res.data(), o, res.size() * 4, cudaMemcpyDeviceToHost);
        int bad = 0;
        for (int c = 0; c < CC; ++c) for (int r2 = 0; r2 < BH; ++r2) for (int i = 0; i < BW; ++i) {
            int gx = x + i, gy = y + r2, gz = z + c;
            float want = (gx >= 0 && gx < W && gy >= 0 && gy < H && gz >= 0 && gz < P) ? (float)(((size_t)gz * H + gy) * W + gx) : 0.f;
            if (res[(c * BH + r2) * BW + i] != want) ++bad;
        }
        printf("mismatches=%d", bad);
    }
    printf("\n");
    return e != cudaSuccess;
}

int main(int argc, char** argv) {
    int which = argc > 1 ? atoi(argv[1]) : 0;
    switch (which) {
        case 0: return run<64, 8, 2>(128, 64, 8, 0, 0, 0);
        case 1: return run<64, 8, 2>(128, 64, 8, 4, 3, 2);
        case 2: return run<64, 8, 2>(128, 64, 8, 5, 3, 2);
        case 3: return run<80, 8, 2>(128, 64, 8, 0, 0, 0);
        case 4: return run<80, 8, 2>(128, 64, 8, 7, 9, 1);
        case 5: return run<80, 8, 2>(128, 64, 8, 100, 60, 7);
        case 6: return run<96, 8, 2>(128, 64, 8, 3, 1, 0);
        case 7: return run<32, 8, 2>(128, 64, 8, 3, 1, 0);
    }
    return 0;
}
