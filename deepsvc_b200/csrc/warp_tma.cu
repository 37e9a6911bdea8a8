// Backward bilinear warp, shared-memory / TMA staged forward (NCHW fp32, sm_100a).
//
// Replaces /root/reference/modules.py:25-62 for the bandwidth-critical calls (the 64-ch
// feature warp of modules.py:429 is 86 % of the hot path's bytes).
//
// One CTA owns a TW x TH tile of output pixels of one batch item:
//   1. the 8 consumer warps compute every pixel's source coordinate (reference
//      arithmetic, warp_common.cuh) and the tile's source bounding box;
//   2. a producer warp streams that box, CC channel planes at a time, from HBM/L2 into
//      a STAGES-deep shared-memory ring with 3-D TMA loads (cp.async.bulk.tensor.3d over
//      the tensor viewed as (x, plane, y): box = BW x CC planes x 8 rows, as many 8-row
//      boxes as the bounding box is tall), signalling mbarriers with complete_tx.  The
//      (x, plane, y) order makes the shared-memory row pitch CC*BW floats = a multiple
//      of 32 banks, so lanes that sample different source rows never bank-conflict;
//   3. the consumers gather the four taps of each of their 8 pixels from shared memory
//      (no tag lookup, no 128-byte-line split: one wavefront per conflict-free LDS
//      instead of two L1 wavefronts per unaligned global gather) and store coalesced
//      128-byte rows.
// Coordinates and weights are computed once per pixel and reused for all C channels.
//
// Work decomposition: grid = (tiles_x, tiles_y, B * csplit).  A CTA stages its tile for
// one of `csplit` channel ranges; the host picks csplit so that the number of CTAs is
// close to a whole number of waves of 2 CTAs x 148 SMs (at 1080p: 1020 tiles = 3.45
// waves -> 2040 half-channel units = 6.9 waves).
//
// A tile whose bounding box does not fit the staging box (wild flow) is cut into
// (32 x 8 pixel) x (8 channel) work items on a device work list (WarpWork).  Every CTA,
// after finishing its own tile, claims items from that list until it is empty, so the
// rare unstageable tiles are gathered by the whole grid inside the same launch (same
// arithmetic, bit-identical results) instead of by one slow CTA or a second launch.
// The last CTA to finish re-zeroes the list: the workspace is zero before and after
// every launch, no memset is needed.
#include <cmath>
#include <cstdlib>
#include <type_traits>
#include <utility>
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
        if (clock64() - t0 > 4000000000ll) __trap();  // ~2 s: protocol bug, do not hang
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

// named barrier over the consumer warps only (the producer warp has returned)
__device__ __forceinline__ void bar_sync_consumers(int nthreads) {
    asm volatile("bar.sync 1, %0;" ::"r"(nthreads) : "memory");
}

template <class F, int... Is>
__device__ __forceinline__ void static_for_impl(F&& f, std::integer_sequence<int, Is...>) {
    (f(std::integral_constant<int, Is>{}), ...);
}
// compile-time loop: the index is usable as a template argument (immediate offsets)
template <int N, class F>
__device__ __forceinline__ void static_for(F&& f) {
    static_for_impl(f, std::make_integer_sequence<int, N>{});
}

template <int IMM>
__device__ __forceinline__ float lds_imm(uint32_t addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(addr), "n"(IMM));
    return v;
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
    static constexpr int ROW_PITCH = CC * BW;                  // smem floats between box rows
    static constexpr int CHUNK_FLOATS = CC * ROWCHUNK * BW;    // one TMA box [8 rows][CC][BW]
    static constexpr int STAGE_FLOATS = (BHMAX / ROWCHUNK) * CHUNK_FLOATS;
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_FLOATS * sizeof(float);
    static_assert(TH % CONSUMER_WARPS == 0 && TW % 32 == 0 && BHMAX % ROWCHUNK == 0, "tile shape");
    static_assert((BW * 4) % 16 == 0 && (CHUNK_FLOATS * 4) % 128 == 0, "TMA alignment");
    static_assert(ROW_PITCH % 32 == 0, "row pitch must be a multiple of the 32 banks");
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
                    WarpParams p, WarpWork* __restrict__ work, int csplit, int cper) {
    constexpr int TW = Cfg::TW, TH = Cfg::TH, BW = Cfg::BW, CC = Cfg::CC, STAGES = Cfg::STAGES;
    constexpr int RPW = Cfg::ROWS_PER_WARP, XH = Cfg::XH, PPT = Cfg::PPT;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage_buf = reinterpret_cast<float*>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ int s_red[Cfg::CONSUMER_WARPS][4];
    __shared__ int s_item;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool is_producer = warp == Cfg::CONSUMER_WARPS;
    const int tx0 = blockIdx.x * TW, ty0 = blockIdx.y * TH;
    const int b = blockIdx.z / csplit, cpart = blockIdx.z - b * csplit;
    const int c_begin = cpart * cper, c_end = min(p.C, c_begin + cper);  // this CTA's channels
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
    // TMA tiled loads need a 16-byte aligned start along the innermost dimension
    // (probed on B200: an unaligned x coordinate raises "illegal instruction")
    const int bx0 = mnx & ~3, by0 = mny;
    const int bw = min(mxx + 1, p.W - 1) - bx0 + 1;
    const int bh = min(mxy + 1, p.H - 1) - mny + 1;
    const bool staged = bw <= BW && bh <= Cfg::BHMAX && c_begin < c_end;  // CTA-uniform
    const int nchunks = (bh + Cfg::ROWCHUNK - 1) / Cfg::ROWCHUNK;
    const int ngroups = (c_end - c_begin + CC - 1) / CC;
    const int plane0 = b * p.C;

    if (is_producer) {
        // ---- phase 2 (producer warp, one elected lane): TMA ring over channel groups.
        // The CTA outlives the loads: the consumers wait on every full barrier.
        if (staged && lane == 0) {
            const uint32_t tx_bytes = (uint32_t)nchunks * Cfg::CHUNK_FLOATS * sizeof(float);
            for (int g = 0; g < ngroups; ++g) {
                const int s = g % STAGES;
                if (g >= STAGES) tma::mbar_wait(&empty_bar[s], ((g / STAGES) - 1) & 1);
                tma::mbar_arrive_expect_tx(&full_bar[s], tx_bytes);
                float* dst = stage_buf + (size_t)s * Cfg::STAGE_FLOATS;
                for (int k = 0; k < nchunks; ++k)
                    tma::load_3d(dst + (size_t)k * Cfg::CHUNK_FLOATS, &tmap, bx0,
                                 plane0 + c_begin + g * CC, by0 + k * Cfg::ROWCHUNK, &full_bar[s]);
            }
        }
        return;
    }

    if (!staged) {
        // unstageable tile: publish it on the work list (once per tile); every CTA of the
        // launch, this one included, helps to gather it in the epilogue below
        if (cpart == 0 && threadIdx.x == 0 && bw > 0 && bh > 0) {
            const int slot = atomicAdd(&work->count, 1);
            const int tile = (b * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x;
            *reinterpret_cast<volatile int*>(&work->items[slot]) = tile + 1;  // non-zero = ready
            __threadfence();
        }
    } else {
        // ---- phase 3 (consumers): gather from the staged box, store coalesced rows.
        // Shared-memory element (row ry, channel c, column rx) of a stage lives at
        //   (ry >> 3) * CHUNK_FLOATS + (ry & 7) * ROW_PITCH + c * BW + rx.
        const bool fast = (tx0 + TW <= p.W) && (ty0 + TH <= p.H) && (mxx + 1 < p.W);  // CTA-uniform
        float* obase = out + (size_t)(plane0 + c_begin) * plane + (size_t)(ty0 + warp * RPW) * p.W + tx0 + lane;
        if (fast) {
            // interior tile, every east tap inside the image: byte addresses, immediate offsets
            float wnw[PPT], wne[PPT], wsw[PPT], wse[PPT];
            uint32_t a_n[PPT], a_s[PPT];
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                const Taps t = make_taps(ixs[k], iys[k], p.W, p.H);
                const int rx = t.x0 - bx0, ry = t.y0 - by0;
                const int ry1 = ry + (t.y1ok ? 1 : 0);
                a_n[k] = 4u * (uint32_t)((ry >> 3) * Cfg::CHUNK_FLOATS + (ry & 7) * Cfg::ROW_PITCH + rx);
                a_s[k] = 4u * (uint32_t)((ry1 >> 3) * Cfg::CHUNK_FLOATS + (ry1 & 7) * Cfg::ROW_PITCH + rx);
                wnw[k] = t.nw;
                wne[k] = t.ne;
                wsw[k] = t.y1ok ? t.sw : 0.0f;
                wse[k] = t.y1ok ? t.se : 0.0f;
            }
            const uint32_t sbase0 = tma::smem_u32(stage_buf);
            for (int g = 0; g < ngroups; ++g) {
                const int s = g % STAGES;
                tma::mbar_wait(&full_bar[s], (g / STAGES) & 1);
                const uint32_t sbase = sbase0 + (uint32_t)s * (Cfg::STAGE_FLOATS * 4);
                uint32_t tn[PPT], ts[PPT];
#pragma unroll
                for (int k = 0; k < PPT; ++k) { tn[k] = a_n[k] + sbase; ts[k] = a_s[k] + sbase; }
                tma::static_for<CC>([&](auto cc) {
                    constexpr int c = decltype(cc)::value;
                    const int ch = g * CC + c;  // relative to c_begin
                    if (c_begin + ch < c_end) {
                        float v[PPT][4];
#pragma unroll
                        for (int k = 0; k < PPT; ++k) {
                            v[k][0] = tma::lds_imm<c * BW * 4>(tn[k]);
                            v[k][1] = tma::lds_imm<c * BW * 4 + 4>(tn[k]);
                            v[k][2] = tma::lds_imm<c * BW * 4>(ts[k]);
                            v[k][3] = tma::lds_imm<c * BW * 4 + 4>(ts[k]);
                        }
                        float* oc = obase + (size_t)ch * plane;
#pragma unroll
                        for (int r = 0; r < RPW; ++r)
#pragma unroll
                            for (int h = 0; h < XH; ++h) {
                                const int k = r * XH + h;
                                float acc = __fmul_rn(v[k][0], wnw[k]);
                                acc = fmaf(v[k][1], wne[k], acc);
                                acc = fmaf(v[k][2], wsw[k], acc);
                                acc = fmaf(v[k][3], wse[k], acc);
                                st_stream1(oc + (size_t)r * p.W + h * 32, acc);
                            }
                    }
                });
                __syncwarp();
                if (lane == 0) tma::mbar_arrive(&empty_bar[s]);
            }
        } else {
            // edge tile (partial, or taps clamped at the right image border): generic loop
            PixelTaps tp[PPT];
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                const Taps t = make_taps(ixs[k], iys[k], p.W, p.H);
                const int rx = t.x0 - bx0, ry = t.y0 - by0;
                const int ry1 = ry + (t.y1ok ? 1 : 0);
                tp[k].off_n = (ry >> 3) * Cfg::CHUNK_FLOATS + (ry & 7) * Cfg::ROW_PITCH + rx;
                tp[k].off_s = (ry1 >> 3) * Cfg::CHUNK_FLOATS + (ry1 & 7) * Cfg::ROW_PITCH + rx;
                tp[k].dx = t.x1ok ? 1 : 0;
                // a tap outside the image contributes nothing (ATen skips it): zero its weight,
                // its (clamped) address stays inside the staged box
                tp[k].nw = t.nw;
                tp[k].ne = t.x1ok ? t.ne : 0.0f;
                tp[k].sw = t.y1ok ? t.sw : 0.0f;
                tp[k].se = (t.x1ok && t.y1ok) ? t.se : 0.0f;
                if (!valid[k]) { tp[k].off_n = tp[k].off_s = 0; tp[k].dx = 0; }
            }
            for (int g = 0; g < ngroups; ++g) {
                const int s = g % STAGES;
                tma::mbar_wait(&full_bar[s], (g / STAGES) & 1);
                const float* sb = stage_buf + (size_t)s * Cfg::STAGE_FLOATS;
#pragma unroll
                for (int c = 0; c < CC; ++c) {
                    const int ch = g * CC + c;
                    if (c_begin + ch < c_end) {
                        const float* sc = sb + c * BW;
                        float* oc = obase + (size_t)ch * plane;
#pragma unroll
                        for (int r = 0; r < RPW; ++r)
#pragma unroll
                            for (int h = 0; h < XH; ++h) {
                                const int k = r * XH + h;
                                const float a = sc[tp[k].off_n], bq = sc[tp[k].off_n + tp[k].dx];
                                const float cq = sc[tp[k].off_s], d = sc[tp[k].off_s + tp[k].dx];
                                float acc = __fmul_rn(a, tp[k].nw);
                                acc = fmaf(bq, tp[k].ne, acc);
                                acc = fmaf(cq, tp[k].sw, acc);
                                acc = fmaf(d, tp[k].se, acc);
                                if (valid[k]) st_stream1(oc + (size_t)r * p.W + h * 32, acc);
                            }
                    }
                }
                __syncwarp();
                if (lane == 0) tma::mbar_arrive(&empty_bar[s]);
            }
        }
    }

    // ---- epilogue (consumers): drain the work list of unstageable tiles, then sign off.
    // Work item n = (tile slot n / per_tile, 32 x 8 pixel block, 8-channel chunk).
    constexpr int ICH = 8;
    const int nchunk_c = (p.C + ICH - 1) / ICH;
    const int per_tile = (TW / 32) * (TH / 8) * nchunk_c;
    for (;;) {
        if (threadIdx.x == 0) {
            int got = -1;
            int n = *reinterpret_cast<volatile int*>(&work->next);
            for (;;) {
                const int avail = *reinterpret_cast<volatile int*>(&work->count) * per_tile;
                if (n >= avail) break;
                const int old = atomicCAS(&work->next, n, n + 1);
                if (old == n) { got = n; break; }
                n = old;
            }
            s_item = got;
        }
        tma::bar_sync_consumers(Cfg::CONSUMER_WARPS * 32);
        const int it = s_item;
        tma::bar_sync_consumers(Cfg::CONSUMER_WARPS * 32);
        if (it < 0) break;
        int tile;
        do {  // the appender stores the tile id right after reserving the slot
            tile = *reinterpret_cast<volatile int*>(&work->items[it / per_tile]);
        } while (tile == 0);
        tile -= 1;
        int r = it % per_tile;
        const int chunk = r % nchunk_c; r /= nchunk_c;
        const int sx = r % (TW / 32), sy = r / (TW / 32);
        const int ttx = tile % gridDim.x, tty = (tile / gridDim.x) % gridDim.y;
        const int tb = tile / (gridDim.x * gridDim.y);
        const int x = ttx * TW + sx * 32 + lane, y = tty * TH + sy * 8 + warp;
        if (x < p.W && y < p.H)
            gather_pixel<ICH>(in, flow, out, lin_x, lin_y, p, tb, x, y, chunk * ICH,
                              min(p.C, chunk * ICH + ICH));
    }
    if (threadIdx.x == 0) {
        __threadfence();
        const int total = gridDim.x * gridDim.y * gridDim.z;
        if (atomicAdd(&work->exited, 1) == total - 1) {
            // last CTA of the launch: leave the workspace zeroed for the next launch
            const int n = *reinterpret_cast<volatile int*>(&work->count);
            for (int i = 0; i < n; ++i) work->items[i] = 0;
            work->count = 0;
            work->next = 0;
            __threadfence();
            work->exited = 0;
        }
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
                      const float* lin_y, const WarpParams& p, void* workspace,
                      size_t workspace_bytes, cudaStream_t st) {
    auto encode = get_encode_fn();
    if (!encode) return -1;
    CUtensorMap tm;
    // tensor viewed as (x, plane, y): the box lands in shared memory as [8 rows][CC][BW]
    const cuuint64_t gdim[3] = {(cuuint64_t)p.W, (cuuint64_t)p.B * p.C, (cuuint64_t)p.H};
    const cuuint64_t gstride[2] = {(cuuint64_t)p.H * p.W * 4, (cuuint64_t)p.W * 4};
    const cuuint32_t box[3] = {(cuuint32_t)Cfg::BW, (cuuint32_t)Cfg::CC, (cuuint32_t)Cfg::ROWCHUNK};
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
    const int tiles_x = (p.W + Cfg::TW - 1) / Cfg::TW, tiles_y = (p.H + Cfg::TH - 1) / Cfg::TH;
    const size_t ntiles = (size_t)tiles_x * tiles_y * p.B;
    if (!workspace || workspace_bytes < sizeof(WarpWork) + ntiles * sizeof(int) || !aligned16(workspace))
        return -1;  // no work list: the caller uses the gather kernel
    // channel split: as close as possible to a whole number of waves of resident CTAs
    static int num_sms = 0, forced_split = -1;
    if (num_sms == 0) {
        int dev = 0;
        if (cudaGetDevice(&dev) != cudaSuccess ||
            cudaDeviceGetAttribute(&num_sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || num_sms <= 0)
            num_sms = DSVC_NUM_SMS;
        const char* e = getenv("DSVC_TMA_CSPLIT");  // tuning knob
        forced_split = e ? atoi(e) : 0;
    }
    const double slots = 2.0 * num_sms;  // __launch_bounds__(THREADS, 2)
    int csplit = 1;
    double best = 0.0;
    for (int cs = 1; cs <= 8; cs *= 2) {
        const int cper = ((p.C + cs - 1) / cs + Cfg::CC - 1) / Cfg::CC * Cfg::CC;
        if (cs > 1 && (cper < 8 || (long long)p.B * cs > 65535)) break;
        const double units = (double)ntiles * cs;
        const double eff = units / (ceil(units / slots) * slots) - 0.01 * (cs - 1);  // flow re-read cost
        if (eff > best + 1e-9) { best = eff; csplit = cs; }
    }
    if (forced_split > 0) csplit = forced_split;
    const int cper = ((p.C + csplit - 1) / csplit + Cfg::CC - 1) / Cfg::CC * Cfg::CC;
    if ((long long)p.B * csplit > 65535) return -1;
    dim3 grid(tiles_x, tiles_y, p.B * csplit);
    warp_fwd_tma_kernel<Cfg><<<grid, Cfg::THREADS, Cfg::SMEM_BYTES, st>>>(
        tm, input, flow, out, lin_x, lin_y, p, static_cast<WarpWork*>(workspace), csplit, cper);
    return (int)cudaGetLastError();
}

// returns -1 when the shape is not eligible (caller uses the gather kernel)
int dsvc_warp_fwd_tma_launch(const float* input, const float* flow, float* out,
                             const float* lin_x, const float* lin_y, const WarpParams& p,
                             bool force, void* workspace, size_t workspace_bytes,
                             cudaStream_t st) {
    // TMA needs 16-byte aligned rows and base; small / few-channel warps gain nothing
    if (p.W % 4 != 0 || !aligned16(input)) return -1;
    if (!force && (p.C < 8 || p.W < 64 || p.H < 32)) return -1;
    if ((long long)p.B * p.C > (1ll << 30)) return -1;
    static int cfg = -1;
    if (cfg < 0) {
        const char* e = getenv("DSVC_TMA_CFG");  // tuning knob
        cfg = e ? atoi(e) : 0;
    }
    switch (cfg) {
        case 1: return launch_cfg<TmaCfg<64, 32, 96, 48, 2, 2>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 2: return launch_cfg<TmaCfg<64, 32, 80, 48, 4, 2>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 3: return launch_cfg<TmaCfg<64, 16, 80, 32, 2, 3>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 4: return launch_cfg<TmaCfg<64, 16, 80, 32, 4, 3>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 5: return launch_cfg<TmaCfg<32, 32, 48, 48, 2, 4>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 6: return launch_cfg<TmaCfg<64, 32, 80, 48, 2, 2>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 7: return launch_cfg<TmaCfg<64, 16, 80, 32, 2, 4>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        default: return launch_cfg<TmaCfg<64, 32, 80, 48, 2, 3>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
    }
}
