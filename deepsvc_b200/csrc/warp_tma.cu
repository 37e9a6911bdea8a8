// Backward bilinear warp, shared-memory / TMA staged forward (NCHW fp32, sm_100a).
//
// Replaces /root/reference/modules.py:25-62 for the bandwidth-critical calls (the 64-ch
// feature warp of modules.py:429 is 86 % of the hot path's bytes).
//
// One CTA owns a TW x TH tile of output pixels of one batch item:
//   1. the 8 consumer warps compute every pixel's source coordinate (reference
//      arithmetic, warp_common.cuh) and the tile's source bounding box;
//   2. a producer warp streams that box, CC channel planes at a time, from HBM/L2 into
//      a STAGES-deep shared-memory ring with 3-D TMA loads
//      (cp.async.bulk.tensor.3d, box = BW x 8 rows x CC planes, as many 8-row boxes as
//      the bounding box is tall), signalling mbarriers with complete_tx;
//   3. the consumers gather the four taps of each of their 8 pixels from shared memory
//      (no tag lookup, no 128-byte-line split: one wavefront per conflict-free LDS
//      instead of two L1 wavefronts per unaligned global gather) and store coalesced
//      128-byte rows.
// Coordinates and weights are computed once per pixel and reused for all C channels.
// A tile whose bounding box does not fit the staging box (wild flow) falls back to the
// read-only-path gather inside the same kernel, so results never depend on the path.
#include <cuda.h>
#include <cudaTypedefs.h>

#include "warp_common.cuh"

namespace dsvc {

namespace tma {

__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void fence_barrier_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
// Parity wait with a wall-clock bound: a protocol bug traps instead of hanging the GPU.
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    const long long t0 = clock64();
    for (;;) {
        uint32_t done;
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        if (clock64() - t0 > 4000000000ll) __trap();
    }
}
__device__ __forceinline__ void load_3d(void* smem_dst, const CUtensorMap* tmap, int x, int y,
                                        int z, uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes "
        "[%0], [%1, {%2, %3, %4}], [%5];" ::"r"(smem_u32(smem_dst)),
        "l"(reinterpret_cast<uint64_t>(tmap)), "r"(x), "r"(y), "r"(z), "r"(smem_u32(bar))
        : "memory");
}

}  // namespace tma

template <int TW_, int TH_, int BW_, int BHMAX_, int CC_, int STAGES_>
struct TmaCfg {
    static constexpr int TW = TW_, TH = TH_, BW = BW_, BHMAX = BHMAX_, CC = CC_, STAGES = STAGES_;
    static constexpr int CONSUMER_WARPS = 8;
    static constexpr int THREADS = (CONSUMER_WARPS + 1) * 32;
    static constexpr int ROWS_PER_WARP = TH / CONSUMER_WARPS;  // rows of the tile per warp
    static constexpr int XH = TW / 32;                         // 32-pixel column groups
    static constexpr int PPT = ROWS_PER_WARP * XH;             // pixels per thread
    static constexpr int ROWCHUNK = 8;                         // rows per TMA box
    static constexpr int CHUNK_FLOATS = CC * ROWCHUNK * BW;    // one TMA box
    static constexpr int STAGE_FLOATS = (BHMAX / ROWCHUNK) * CHUNK_FLOATS;
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_FLOATS * sizeof(float);
    static_assert(TH % CONSUMER_WARPS == 0 && TW % 32 == 0 && BHMAX % ROWCHUNK == 0, "tile shape");
    static_assert((BW * 4) % 16 == 0 && (CHUNK_FLOATS * 4) % 128 == 0, "TMA alignment");
};

struct PixelTaps {
    float nw, ne, sw, se;
    int off_n, off_s;  // shared-memory offsets of the north / south tap rows (channel 0)
    int dx;            // 1 if the east taps are inside the image, else 0
};

template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 2)
warp_fwd_tma_kernel(const __grid_constant__ CUtensorMap tmap, const float* __restrict__ in,
                    const float* __restrict__ flow, float* __restrict__ out,
                    const float* __restrict__ lin_x, const float* __restrict__ lin_y,
                    WarpParams p) {
    constexpr int TW = Cfg::TW, TH = Cfg::TH, BW = Cfg::BW, CC = Cfg::CC, STAGES = Cfg::STAGES;
    constexpr int RPW = Cfg::ROWS_PER_WARP, XH = Cfg::XH, PPT = Cfg::PPT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage_buf = reinterpret_cast<float*>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ int s_red[Cfg::CONSUMER_WARPS][4];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool is_producer = warp == Cfg::CONSUMER_WARPS;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH, b = blockIdx.z;
    const size_t plane = (size_t)p.H * p.W;

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            tma::mbar_init(&full_bar[s], 1);
            tma::mbar_init(&empty_bar[s], Cfg::CONSUMER_WARPS);
        }
        tma::fence_barrier_init();
    }

    // ---- phase 1: per-pixel coordinates (consumers) and the tile's source bounding box
    float ixs[PPT], iys[PPT];
    int x0s[PPT], y0s[PPT];
    bool valid[PPT];
    int mnx = INT_MAX, mxx = INT_MIN, mny = INT_MAX, mxy = INT_MIN;
    if (!is_producer) {
        const float* fl = flow + (size_t)b * 2 * plane;
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
#pragma unroll
            for (int h = 0; h < XH; ++h) {
                const int k = r * XH + h;
                const int x = tx0 + h * 32 + lane, y = ty0 + warp * RPW + r;
                valid[k] = x < p.W && y < p.H;
                ixs[k] = iys[k] = 0.0f;
                x0s[k] = y0s[k] = 0;
                if (valid[k]) {
                    const size_t pix = (size_t)y * p.W + x;
                    const float fx = __ldg(fl + pix), fy = __ldg(fl + plane + pix);
                    ixs[k] = source_coord(__ldg(lin_x + x), fx, p.sx, p.inv_sx, p.flow_mode, p.W);
                    iys[k] = source_coord(__ldg(lin_y + y), fy, p.sy, p.inv_sy, p.flow_mode, p.H);
                    x0s[k] = (int)floorf(ixs[k]);
                    y0s[k] = (int)floorf(iys[k]);
                    mnx = min(mnx, x0s[k]); mxx = max(mxx, x0s[k]);
                    mny = min(mny, y0s[k]); mxy = max(mxy, y0s[k]);
                }
            }
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) {
            mnx = min(mnx, __shfl_xor_sync(0xffffffffu, mnx, o));
            mxx = max(mxx, __shfl_xor_sync(0xffffffffu, mxx, o));
            mny = min(mny, __shfl_xor_sync(0xffffffffu, mny, o));
            mxy = max(mxy, __shfl_xor_sync(0xffffffffu, mxy, o));
        }
        if (lane == 0) {
            s_red[warp][0] = mnx; s_red[warp][1] = mxx; s_red[warp][2] = mny; s_red[warp][3] = mxy;
        }
    }
    __syncthreads();  // also publishes the barrier initialisation
    mnx = INT_MAX; mxx = INT_MIN; mny = INT_MAX; mxy = INT_MIN;
#pragma unroll
    for (int w = 0; w < Cfg::CONSUMER_WARPS; ++w) {
        mnx = min(mnx, s_red[w][0]); mxx = max(mxx, s_red[w][1]);
        mny = min(mny, s_red[w][2]); mxy = max(mxy, s_red[w][3]);
    }
    // taps reach x0+1 / y0+1 (clamped to the image)
    const int bx0 = mnx, by0 = mny;
    const int bw = min(mxx + 1, p.W - 1) - mnx + 1;
    const int bh = min(mxy + 1, p.H - 1) - mny + 1;
    const bool staged = bw <= BW && bh <= Cfg::BHMAX;  // CTA-uniform
    const int nchunks = (bh + Cfg::ROWCHUNK - 1) / Cfg::ROWCHUNK;
    const int ngroups = (p.C + CC - 1) / CC;
    const int plane0 = b * p.C;

    if (!staged) {
        // ---- fallback: direct gather of this tile (identical arithmetic)
        if (is_producer) return;
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
#pragma unroll
            for (int h = 0; h < XH; ++h) {
                const int k = r * XH + h;
                if (!valid[k]) continue;
                const int x = tx0 + h * 32 + lane, y = ty0 + warp * RPW + r;
                const Taps t = make_taps(ixs[k], iys[k], p.W, p.H);
                const int dx = t.x1ok ? 1 : 0, dy = t.y1ok ? p.W : 0;
                const float* ip = in + (size_t)plane0 * plane + (size_t)t.y0 * p.W + t.x0;
                float* op = out + (size_t)plane0 * plane + (size_t)y * p.W + x;
#pragma unroll 4
                for (int c = 0; c < p.C; ++c) {
                    float acc = __fmul_rn(__ldg(ip), t.nw);
                    acc = t.x1ok ? fmaf(__ldg(ip + dx), t.ne, acc) : acc;
                    acc = t.y1ok ? fmaf(__ldg(ip + dy), t.sw, acc) : acc;
                    acc = (t.x1ok && t.y1ok) ? fmaf(__ldg(ip + dy + dx), t.se, acc) : acc;
                    st_stream1(op, acc);
                    ip += plane;
                    op += plane;
                }
            }
        }
        return;
    }

    if (is_producer) {
        // ---- phase 2 (producer warp, one elected lane): TMA ring over channel groups
        if (lane == 0) {
            const uint32_t tx_bytes = (uint32_t)nchunks * Cfg::CHUNK_FLOATS * sizeof(float);
            for (int g = 0; g < ngroups; ++g) {
                const int s = g % STAGES;
                if (g >= STAGES) tma::mbar_wait(&empty_bar[s], ((g / STAGES) - 1) & 1);
                tma::mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
                float* dst = stage_buf + (size_t)s * Cfg::STAGE_FLOATS;
                for (int k = 0; k < nchunks; ++k)
                    tma::load_3d(dst + (size_t)k * Cfg::CHUNK_FLOATS, &tmap, bx0,
                                 by0 + k * Cfg::ROWCHUNK, plane0 + g * CC, &full_bar[s]);
            }
        }
        return;
    }

    // ---- phase 3 (consumers): gather from the staged box, store coalesced rows
    PixelTaps tp[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        const Taps t = make_taps(ixs[k], iys[k], p.W, p.H);
        const int rx = t.x0 - bx0, ry = t.y0 - by0;
        const int ry1 = ry + (t.y1ok ? 1 : 0);
        tp[k].off_n = (ry >> 3) * Cfg::CHUNK_FLOATS + (ry & 7) * BW + rx;
        tp[k].off_s = (ry1 >> 3) * Cfg::CHUNK_FLOATS + (ry1 & 7) * BW + rx;
        tp[k].dx = t.x1ok ? 1 : 0;
        // a tap outside the image contributes nothing (ATen skips it): zero its weight,
        // its (clamped) address stays inside the staged box
        tp[k].nw = t.nw;
        tp[k].ne = t.x1ok ? t.ne : 0.0f;
        tp[k].sw = t.y1ok ? t.sw : 0.0f;
        tp[k].se = (t.x1ok && t.y1ok) ? t.se : 0.0f;
        if (!valid[k]) { tp[k].off_n = tp[k].off_s = 0; tp[k].dx = 0; }
    }
    float* obase = out + (size_t)plane0 * plane + (size_t)(ty0 + warp * RPW) * p.W + tx0 + lane;
    for (int g = 0; g < ngroups; ++g) {
        const int s = g % STAGES;
        tma::mbar_wait(&full_bar[s], (g / STAGES) & 1);
        const float* sb = stage_buf + (size_t)s * Cfg::STAGE_FLOATS;
#pragma unroll
        for (int c = 0; c < CC; ++c) {
            const int ch = g * CC + c;
            if (ch < p.C) {
                const float* sc = sb + c * (Cfg::ROWCHUNK * BW);
                float* oc = obase + (size_t)ch * plane;
                float res[PPT];
#pragma unroll
                for (int k = 0; k < PPT; ++k) {
                    const float a = sc[tp[k].off_n], bq = sc[tp[k].off_n + tp[k].dx];
                    const float cq = sc[tp[k].off_s], d = sc[tp[k].off_s + tp[k].dx];
                    float acc = __fmul_rn(a, tp[k].nw);
                    acc = fmaf(bq, tp[k].ne, acc);
                    acc = fmaf(cq, tp[k].sw, acc);
                    res[k] = fmaf(d, tp[k].se, acc);
                }
#pragma unroll
                for (int r = 0; r < RPW; ++r)
#pragma unroll
                    for (int h = 0; h < XH; ++h)
                        if (valid[r * XH + h]) st_stream1(oc + (size_t)r * p.W + h * 32, res[r * XH + h]);
            }
        }
        __syncwarp();
        if (lane == 0) tma::mbar_arrive(&empty_bar[s]);
    }
}

}  // namespace dsvc

using namespace dsvc;

// ------------------------------------------------------------------------ host side
static PFN_cuTensorMapEncodeTiled_v12000 get_encode_fn() {
    static PFN_cuTensorMapEncodeTiled_v12000 fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(p);
    }
    return fn;
}

template <class Cfg>
static int launch_cfg(const float* input, const float* flow, float* out, const float* lin_x,
                      const float* lin_y, const WarpParams& p, cudaStream_t st) {
    auto encode = get_encode_fn();
    if (!encode) return -1;
    CUtensorMap tm;
    const cuuint64_t gdim[3] = {(cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B * p.C};
    const cuuint64_t gstride[2] = {(cuuint64_t)p.W * 4, (cuuint64_t)p.H * p.W * 4};
    const cuuint32_t box[3] = {(cuuint32_t)Cfg::BW, (cuuint32_t)Cfg::ROWCHUNK, (cuuint32_t)Cfg::CC};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(input),
                              gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                              CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -1;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(warp_fwd_tma_kernel<Cfg>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        attr_set = true;
    }
    dim3 grid((p.W + Cfg::TW - 1) / Cfg::TW, (p.H + Cfg::TH - 1) / Cfg::TH, p.B);
    warp_fwd_tma_kernel<Cfg><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(tm, input, flow, out,
                                                                          lin_x, lin_y, p);
    return (int)cudaGetLastError();
}

// returns -1 when the shape is not eligible (caller uses the gather kernel)
int dsvc_warp_fwd_tma_launch(const float* input, const float* flow, float* out,
                             const float* lin_x, const float* lin_y, const WarpParams& p,
                             bool force, cudaStream_t st) {
    // TMA needs 16-byte aligned rows and base; small / few-channel warps gain nothing
    if (p.W % 4 != 0 || !aligned16(input)) return -1;
    if (!force && (p.C < 8 || p.W < 64 || p.H < 32)) return -1;
    if ((long long)p.B * p.C > (1ll << 30)) return -1;
    using Cfg = TmaCfg<64, 32, 80, 48, 2, 3>;
    return launch_cfg<Cfg>(input, flow, out, lin_x, lin_y, p, st);
}
