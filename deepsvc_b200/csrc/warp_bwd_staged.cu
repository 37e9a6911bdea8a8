// Backward of the bilinear warp, shared-memory staged (NCHW fp32, sm_100a).
//
// Replaces the autograd of /root/reference/modules.py:25-62 (ATen grid_sampler_2d_backward)
// for the bandwidth-critical calls: the 64-ch feature warp of modules.py:429 in training
// (Learner.py:1343) needs grad_input and grad_flow.
//
// Why not scatter straight to global memory: the four RED.ADD.F32 per output element of the
// per-pixel kernel reach L2 as ~2 useful adds per 32-byte sector (jittered flows spread a
// warp's taps over two rows and repeat columns), and L2 processes reductions per sector --
// r01 ncu: 249 M reduction sectors, L2 reduction pipe 56 % busy, 1.25 TB/s of algorithmic bytes.
//
// This kernel combines the scatter per 64 x 16 output tile before it leaves the SM, without
// shared-memory float atomics (a CAS loop on sm_100a):
//   * once per tile (the geometry is the same for every channel) the CTA inverts the
//     scatter: it counts the taps landing on every element of the tile's source bounding box
//     (integer shared-memory atomics), prefix-sums the counts and writes the (weight, source
//     pixel) pairs in destination order -- a CSR matrix of the transposed warp;
//   * every thread then keeps 16 consecutive pairs in registers.  Per channel it gathers
//     grad_out of its pairs from the TMA-staged grad_out tile, sums each destination
//     element's run in a fixed order and writes it once to a shared-memory out-box (an element
//     whose pairs cross a thread boundary is summed piecewise: the pieces meet in the zeroed
//     out-box through shared-memory atomics, at most 2 per thread and channel); the CTA then
//     adds the touched 16-byte quads of the out-box to grad_input with row-contiguous
//     RED.ADD.v4.F32 (full 32-byte sectors, the halo shared with neighbouring tiles included;
//     scripts/probe/red_probe.cu: 1.3 T element-adds/s against 0.65 T for
//     cp.reduce.async.bulk from the same box);
//   * grad_flow: the four input taps of every pixel come from the TMA-staged input box (as
//     in the forward kernel), reduced over the channels in registers.
// Input box and grad_out tile are pipelined over the channels through an mbarrier ring; the
// out-box is double-buffered (the adds of channel i are issued while channel i + 1 is summed).
// A tile whose bounding box does not fit the staging box takes the per-pixel direct scatter
// of warp_bwd_common.cuh inside the same launch.
#include <algorithm>
#include <climits>
#include <cstdlib>

#include "tma_utils.cuh"
#include "warp_bwd_common.cuh"

namespace dsvc {

namespace bwd {
constexpr int TW = 64, TH = 16;          // output tile
constexpr int BW = 96, BH = 32;          // staged source box (floats x rows), as the forward kernel's
constexpr int ROWCHUNK = 8;              // rows per TMA box
constexpr int NE = BW * BH;              // destination elements of a box
constexpr int THREADS = 256;
constexpr int PPT = TW * TH / THREADS;   // pixels per thread (4)
constexpr int NS = 4;                    // load stages
constexpr int SHARE = TW * TH * 4 / THREADS;  // (weight, source pixel) pairs per thread (16)
constexpr int KMAX = SHARE;              // register pair slots per thread
constexpr int GB = 4;                    // gathers issued back to back
constexpr int QPT = (BH * (BW / 4) + THREADS - 1) / THREADS;  // out-box quads per thread (2)
constexpr int STAGE_IN = NE, STAGE_G = TW * TH;
constexpr int STAGE_FLOATS = STAGE_IN + STAGE_G;
constexpr int OB_FLOATS = NE + 3 * THREADS;  // out-box + per-thread private slots (head piece, tail piece, sink)
constexpr size_t SMEM_BYTES = (size_t)(NS * STAGE_FLOATS + 2 * OB_FLOATS + NE) * 4;
static_assert(NS % 2 == 0, "the out-box parity follows the stage parity");
static_assert(NS * STAGE_FLOATS * 4 >= TW * TH * 4 * 8, "the pair scratch aliases the load stages");
static_assert((STAGE_FLOATS * 4) % 128 == 0 && (STAGE_IN * 4) % 128 == 0 && (ROWCHUNK * BW * 4) % 128 == 0, "TMA alignment");
constexpr int EPT = NE / THREADS;         // scan: destination elements per thread (12)
static_assert(NE % THREADS == 0 && EPT % 4 == 0, "scan: whole int4 per thread");

__device__ __forceinline__ void red_add4(float* p, float4 v) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float4 lds4(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a));
    return v;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }
__device__ __forceinline__ float lds(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void red_shared(uint32_t a, float v) {  // CAS loop on sm_100a: rare pieces only
    asm volatile("red.shared.add.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
__device__ __forceinline__ void sts4(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(a), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
template <int IMM>
__device__ __forceinline__ float lds_i(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(IMM));
    return v;
}
template <int IMM>
__device__ __forceinline__ void sts_i(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0+%1], %2;" ::"r"(a), "n"(IMM), "f"(v) : "memory");
}
template <int IMM>
__device__ __forceinline__ float4 lds4_i(uint32_t a) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4+%5];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(a), "n"(IMM));
    return v;
}
template <int IMM>
__device__ __forceinline__ void sts4_i(uint32_t a, float4 v) {
    asm volatile("st.shared.v4.f32 [%0+%1], {%2, %3, %4, %5};" ::"r"(a), "n"(IMM), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ void sts(uint32_t a, float v) {
    asm volatile("st.shared.f32 [%0], %1;" ::"r"(a), "f"(v) : "memory");
}
}  // namespace bwd

// A tile whose taps do not fit the staging box (wild flow): direct scatter, channel-outer so
// that the taps of all the thread's pixels are in flight together.  Not inlined: the staged
// path keeps its registers (the coordinates are recomputed here).
template <bool NEED_GIN>
__device__ __noinline__ void bwd_tile_direct(const float* __restrict__ gout, const float* __restrict__ in,
                                             const float* __restrict__ flow, float* __restrict__ gin,
                                             float* __restrict__ gflow, const float* __restrict__ lin_x,
                                             const float* __restrict__ lin_y, const WarpParams& p, int b,
                                             int tx0, int ty0, int c_begin, int c_end, bool acc_gflow) {
    using namespace bwd;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t plane = (size_t)p.H * p.W;
    BwdCoord bc[PPT];
    bool valid[PPT];
    int o_nw[PPT], dxg[PPT], dyg[PPT];
    float gx[PPT], gy[PPT];
    const float* fl = flow + (size_t)b * 2 * plane;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        const int x = tx0 + (k & 1) * 32 + lane, y = ty0 + warp * 2 + (k >> 1);
        valid[k] = x < p.W && y < p.H;
        const int xc = min(x, p.W - 1), yc = min(y, p.H - 1);
        const size_t pix = (size_t)yc * p.W + xc;
        bc[k] = bwd_coord(__ldg(lin_x + xc), __ldg(lin_y + yc), __ldg(fl + pix), __ldg(fl + plane + pix), p);
        o_nw[k] = bc[k].t.y0 * p.W + bc[k].t.x0;
        dxg[k] = bc[k].t.x1ok ? 1 : 0;
        dyg[k] = bc[k].t.y1ok ? p.W : 0;
        gx[k] = gy[k] = 0.0f;
    }
    const size_t pix0 = (size_t)(ty0 + warp * 2) * p.W + tx0 + lane;
    for (int c = c_begin; c < c_end; ++c) {
        const size_t cb = (size_t)(b * p.C + c) * plane;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            if (!valid[k]) continue;
            const float g = __ldg(gout + cb + pix0 + (size_t)(k >> 1) * p.W + (k & 1) * 32);
            if (NEED_GIN) {
                float* gi = gin + cb + o_nw[k];
                atomicAdd(gi, __fmul_rn(bc[k].t.nw, g));
                if (dxg[k]) atomicAdd(gi + 1, __fmul_rn(bc[k].t.ne, g));
                if (dyg[k]) atomicAdd(gi + dyg[k], __fmul_rn(bc[k].t.sw, g));
                if (dxg[k] && dyg[k]) atomicAdd(gi + dyg[k] + 1, __fmul_rn(bc[k].t.se, g));
            }
            if (gflow) {
                const float* ip = in + cb + o_nw[k];
                const float v_nw = __ldg(ip), v_ne = __ldg(ip + dxg[k]);
                const float v_sw = __ldg(ip + dyg[k]), v_se = __ldg(ip + dyg[k] + dxg[k]);
                const float tx = fmaf(bc[k].wy1, v_se - v_sw, bc[k].wy0 * (v_ne - v_nw));
                const float ty = fmaf(bc[k].wx1, v_se - v_ne, bc[k].wx0 * (v_sw - v_nw));
                gx[k] = fmaf(tx, g, gx[k]);
                gy[k] = fmaf(ty, g, gy[k]);
            }
        }
    }
    if (gflow) {
#pragma unroll
        for (int k = 0; k < PPT; ++k)
            if (valid[k])
                store_gflow(gflow, p, b, pix0 + (size_t)(k >> 1) * p.W + (k & 1) * 32, bc[k], gx[k], gy[k], acc_gflow);
    }
}

// grid.x = tiles * csplit; unit u -> tile u / csplit, channel range (u % csplit) * cper ...
template <bool NEED_GIN>
__global__ void __launch_bounds__(bwd::THREADS, 2)
warp_bwd_staged_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_gout,
                       const float* __restrict__ gout,
                       const float* __restrict__ in, const float* __restrict__ flow,
                       float* __restrict__ gin, float* __restrict__ gflow,
                       const float* __restrict__ lin_x, const float* __restrict__ lin_y, WarpParams p,
                       int tiles_x, int tiles_y, int csplit, int cper, int perm_mul,
                       const int* __restrict__ mode) {
    using namespace bwd;
    // (launch-level fallback, see dsvc_warp_bwd_ws_f32: the scout launch decided that the per-pixel
    // kernel, enqueued right behind this one, takes the whole job)
    if (mode && *mode == DSVC_BWD_MODE_DIRECT) return;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stages = reinterpret_cast<float*>(smem_raw);          // NS x [in box | grad_out tile]
    float* outbox = stages + NS * STAGE_FLOATS;                  // 2 x ([BH][BW] + private slots)
    int* cursor = reinterpret_cast<int*>(outbox + 2 * OB_FLOATS);  // [NE] tap counts -> first-pair offsets
    uint2* pairs = reinterpret_cast<uint2*>(stages);             // [<= 4096] (weight, e << 10 | q), scratch
    __shared__ __align__(8) uint64_t full_bar[NS];
    __shared__ int red_i[8][4];
    __shared__ int scan_w[8];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int unit = blockIdx.x;
    const int tile = unit / csplit, cs = unit - tile * csplit;
    const int c_begin = cs * cper, c_end = min(p.C, c_begin + cper);
    if (c_begin >= c_end) return;
    const int tx0 = (tile % tiles_x) * TW, ty0 = ((tile / tiles_x) % tiles_y) * TH;
    const int b = tile / (tiles_x * tiles_y);
    const size_t plane = (size_t)p.H * p.W;
    const bool acc_gflow = csplit > 1;

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) tma::mbar_init(&full_bar[s], 1);
        tma::fence_barrier_init();
    }

    // ---- geometry of the thread's pixels: k = r * 2 + h -> (xx, yy) = (h * 32 + lane, warp * 2 + r)
    BwdCoord bc[PPT];
    bool valid[PPT];
    int mnx = INT_MAX, mxx = INT_MIN, mny = INT_MAX, mxy = INT_MIN;
    {
        const float* fl = flow + (size_t)b * 2 * plane;
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            const int xx = (k & 1) * 32 + lane, yy = warp * 2 + (k >> 1);
            const int x = tx0 + xx, y = ty0 + yy;
            valid[k] = x < p.W && y < p.H;
            const int xc = min(x, p.W - 1), yc = min(y, p.H - 1);
            const size_t pix = (size_t)yc * p.W + xc;
            bc[k] = bwd_coord(__ldg(lin_x + xc), __ldg(lin_y + yc), __ldg(fl + pix), __ldg(fl + plane + pix), p);
            if (valid[k]) {
                mnx = min(mnx, bc[k].t.x0); mxx = max(mxx, bc[k].t.x0);
                mny = min(mny, bc[k].t.y0); mxy = max(mxy, bc[k].t.y0);
            }
        }
    }
    mnx = __reduce_min_sync(0xffffffffu, mnx); mxx = __reduce_max_sync(0xffffffffu, mxx);
    mny = __reduce_min_sync(0xffffffffu, mny); mxy = __reduce_max_sync(0xffffffffu, mxy);
    if (lane == 0) { red_i[warp][0] = mnx; red_i[warp][1] = mxx; red_i[warp][2] = mny; red_i[warp][3] = mxy; }
    // zero the tap counts and both out-boxes (elements no tap lands on stay zero for every channel)
    for (int i = tid; i < NE / 4; i += THREADS) {
        reinterpret_cast<int4*>(cursor)[i] = make_int4(0, 0, 0, 0);
        reinterpret_cast<float4*>(outbox)[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
        reinterpret_cast<float4*>(outbox + OB_FLOATS)[i] = make_float4(0.0f, 0.0f, 0.0f, 0.0f);
    }
    __syncthreads();
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        mnx = min(mnx, red_i[w][0]); mxx = max(mxx, red_i[w][1]);
        mny = min(mny, red_i[w][2]); mxy = max(mxy, red_i[w][3]);
    }
    const int bx0 = mnx & ~3, by0 = mny;  // TMA boxes start 16-byte aligned along x
    const int bw = min(mxx + 1, p.W - 1) - bx0 + 1, bh = min(mxy + 1, p.H - 1) - by0 + 1;
    bool staged = mnx <= mxx && bw <= BW && bh <= BH;

    // element (ry * BW + rx) of the box that pixel k's north-west tap lands on, east / south steps
    int e_nw[PPT], dxs[PPT], dys[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        e_nw[k] = (bc[k].t.y0 - by0) * BW + (bc[k].t.x0 - bx0);
        dxs[k] = bc[k].t.x1ok ? 1 : 0;
        dys[k] = bc[k].t.y1ok ? BW : 0;
    }

    // ---- transposed warp of the tile as CSR (grad_input only)
    float pw[KMAX];        // weight (0 for an empty slot)
    uint32_t pa[KMAX];     // shared-window address of the source pixel in stage 0's grad_out tile
    uint32_t pd[KMAX];     // shared-window address the running sum is stored to after this pair: the
                           // destination element in out-box 0 if the pair ends a whole run, the thread's
                           // head / tail slot if it ends a piece of a split run, else the thread's sink
    uint32_t lastmask = 0; // bit j: pair j ends a run (the running sum restarts); bit 16 / 17: head / tail piece
    uint32_t head_dst = 0, tail_dst = 0;  // out-box addresses the pieces are added to
    int npairs = 0;
    uint32_t quad_s[QPT], quad_g[QPT];  // out-box byte offset (or ~0) / element offset in the plane
    if (NEED_GIN && staged) {
        int rank[PPT][4];
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            if (!valid[k]) continue;
            // the returned count is the tap's rank inside its destination element: kept, so that
            // the fill below needs no second pass of atomics (integer ATOMS run at 2 cycles per lane)
            rank[k][0] = atomicAdd(&cursor[e_nw[k]], 1);
            if (dxs[k]) rank[k][1] = atomicAdd(&cursor[e_nw[k] + 1], 1);
            if (dys[k]) rank[k][2] = atomicAdd(&cursor[e_nw[k] + BW], 1);
            if (dxs[k] && dys[k]) rank[k][3] = atomicAdd(&cursor[e_nw[k] + BW + 1], 1);
        }
        __syncthreads();
        // the thread's 16-byte quads of the out-box (<= QPT of the bh x ceil(bw / 4) the box spans)
        // that at least one tap lands on: added to grad_input after every channel
        {
            const int nq = (bw + 3) >> 2;
#pragma unroll
            for (int u = 0; u < QPT; ++u) {
                const int idx = tid + u * THREADS;
                const int r = idx / nq, q = idx - r * nq;
                quad_s[u] = 0xffffffffu;
                quad_g[u] = 0u;
                if (r < bh) {
                    const int4 c4 = *reinterpret_cast<const int4*>(cursor + r * BW + 4 * q);
                    if (c4.x | c4.y | c4.z | c4.w) {
                        quad_s[u] = 4u * (uint32_t)(r * BW + 4 * q);
                        quad_g[u] = (uint32_t)((by0 + r) * p.W + bx0 + 4 * q);
                    }
                }
            }
        }
        // exclusive scan of the counts (EPT elements per thread)
        int cnt[EPT], sum = 0;
#pragma unroll
        for (int v = 0; v < EPT / 4; ++v) {
            const int4 a = reinterpret_cast<const int4*>(cursor)[tid * (EPT / 4) + v];
            cnt[4 * v] = a.x; cnt[4 * v + 1] = a.y; cnt[4 * v + 2] = a.z; cnt[4 * v + 3] = a.w;
        }
#pragma unroll
        for (int j = 0; j < EPT; ++j) sum += cnt[j];
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) scan_w[warp] = incl;
        __syncthreads();
        int base = 0, total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            base += w < warp ? scan_w[w] : 0;
            total += scan_w[w];
        }
        if (staged) {
            int run = base + incl - sum;
#pragma unroll
            for (int v = 0; v < EPT / 4; ++v) {
                int o[4];
#pragma unroll
                for (int j = 0; j < 4; ++j) { o[j] = run; run += cnt[4 * v + j]; }
                reinterpret_cast<int4*>(cursor)[tid * (EPT / 4) + v] = make_int4(o[0], o[1], o[2], o[3]);
            }
            __syncthreads();
            // fill: pairs land in destination order (inside one destination in the order the counting
            // atomics were served)
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                if (!valid[k]) continue;
                const uint32_t q = (uint32_t)((warp * 2 + (k >> 1)) * TW + (k & 1) * 32 + lane);
                auto put = [&](int e, float w, int r) {  // cursor[e] = first pair of element e
                    pairs[cursor[e] + r] = make_uint2(__float_as_uint(w), ((uint32_t)e << 10) | q);
                };
                put(e_nw[k], bc[k].t.nw, rank[k][0]);
                if (dxs[k]) put(e_nw[k] + 1, bc[k].t.ne, rank[k][1]);
                if (dys[k]) put(e_nw[k] + BW, bc[k].t.sw, rank[k][2]);
                if (dxs[k] && dys[k]) put(e_nw[k] + BW + 1, bc[k].t.se, rank[k][3]);
            }
            __syncthreads();
            // the thread's share: 16 consecutive pairs (lane l of warp w takes share 32 w + (m l mod 32);
            // m = 1: other odd multipliers were measured and change nothing).  A destination element
            // whose run of pairs crosses a share boundary is summed piecewise: the pieces go to the
            // thread's private head / tail slot and are added to the (zeroed) out-box with
            // shared-memory atomics after the loop; whole runs are stored.
            const int share = warp * 32 + ((lane * perm_mul) & 31);
            const int s = SHARE * share, e = min(SHARE * (share + 1), total);
            npairs = max(e - s, 0);
            uint32_t el_before = 0xffffffffu, el_after = 0xffffffffu, el_first = 0u;
            if (npairs > 0) {
                if (s > 0) el_before = pairs[s - 1].y >> 10;
                if (e < total) el_after = pairs[e].y >> 10;
                el_first = pairs[s].y >> 10;
            }
            const uint32_t ob0 = tma::smem_u32(outbox), gt0 = tma::smem_u32(stages) + (uint32_t)(STAGE_IN * 4);
            const uint32_t slot_head = ob0 + 4u * (uint32_t)(NE + tid), slot_tail = slot_head + 4u * THREADS,
                           slot_sink = slot_tail + 4u * THREADS;
#pragma unroll
            for (int j = 0; j < KMAX; ++j) {
                pw[j] = 0.0f;
                pa[j] = gt0;
                pd[j] = slot_sink;
                if (j < npairs) {
                    const uint2 pr = pairs[s + j];
                    const uint32_t el = pr.y >> 10;
                    const bool last = (j == npairs - 1) || (pairs[s + j + 1].y >> 10) != el;
                    pw[j] = __uint_as_float(pr.x);
                    pa[j] = gt0 + ((pr.y & 1023u) << 2);
                    if (last) {
                        lastmask |= 1u << j;
                        if (el == el_first && el == el_before) {  // continues the previous share's run
                            pd[j] = slot_head;
                            head_dst = ob0 + 4u * el;
                            lastmask |= 1u << 16;
                        } else if (el == el_after) {               // continued by the next share
                            pd[j] = slot_tail;
                            tail_dst = ob0 + 4u * el;
                            lastmask |= 1u << 17;
                        } else {
                            pd[j] = ob0 + 4u * el;
                        }
                    }
                }
            }
        }
        __syncthreads();  // the pair scratch is dead: the load stages may be overwritten
    }

    if (!staged) {  // CTA-uniform
        bwd_tile_direct<NEED_GIN>(gout, in, flow, gin, gflow, lin_x, lin_y, p, b, tx0, ty0, c_begin, c_end, acc_gflow);
        return;
    }

    // ---- channel loop
    const int nch = c_end - c_begin;
    const int nchunks = gflow ? (bh + ROWCHUNK - 1) / ROWCHUNK : 0;  // the input box feeds grad_flow only
    const uint32_t sbase0 = tma::smem_u32(stages), obase0 = tma::smem_u32(outbox), full0 = tma::smem_u32(full_bar);
    const uint32_t tx_bytes = (uint32_t)(nchunks * ROWCHUNK * BW + STAGE_G) * 4u;
    const int plane0 = b * p.C + c_begin;
    auto issue_loads = [&](int i) {  // channel i of the range into stage i % NS (one thread)
        const uint32_t s = (uint32_t)(i % NS);
        const uint32_t dst = sbase0 + s * (uint32_t)(STAGE_FLOATS * 4), bar = full0 + 8u * s;
        tma::mbar_arrive_expect_tx(bar, tx_bytes);
        for (int k = 0; k < nchunks; ++k)
            tma::load_3d(dst + (uint32_t)(k * ROWCHUNK * BW * 4), &tm_in, bx0, by0 + k * ROWCHUNK, plane0 + i, bar);
        tma::load_3d(dst + (uint32_t)(STAGE_IN * 4), &tm_gout, tx0, ty0, plane0 + i, bar);
    };
    if (tid == 0) {
        fence_async_smem();  // generic-proxy writes to the stages (pair scratch) before the async-proxy loads
        for (int i = 0; i < NS && i < nch; ++i) issue_loads(i);
    }

    // byte offsets of the taps inside a stage, weights for grad_flow
    uint32_t a_n[PPT], a_s[PPT], dx4[PPT];
    float w1[PPT][4];  // wx0, wx1, wy0, wy1
    float gix[PPT], giy[PPT];
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        a_n[k] = 4u * (uint32_t)(valid[k] ? e_nw[k] : 0);
        a_s[k] = 4u * (uint32_t)(valid[k] ? e_nw[k] + dys[k] : 0);
        dx4[k] = valid[k] ? 4u * (uint32_t)dxs[k] : 0u;
        w1[k][0] = bc[k].wx0; w1[k][1] = bc[k].wx1; w1[k][2] = bc[k].wy0; w1[k][3] = bc[k].wy1;
        gix[k] = giy[k] = 0.0f;
    }
    // absolute shared-window addresses in stage 0 / out-box 0; stage and buffer offsets are immediates
#pragma unroll
    for (int k = 0; k < PPT; ++k) { a_n[k] += sbase0; a_s[k] += sbase0; }
    // the thread's own pixels in the grad_out tile: one base address, immediates per pixel
    const uint32_t a_g0 = sbase0 + 4u * (uint32_t)(STAGE_IN + warp * 2 * TW + lane);
#pragma unroll
    for (int u = 0; u < QPT; ++u) quad_s[u] += quad_s[u] != 0xffffffffu ? obase0 : 0u;

    const uint32_t ob_head = obase0 + 4u * (uint32_t)(NE + tid);
    // one channel: stage ST (compile time), out-box ST & 1
    auto channel = [&](auto stc, int i) {
        constexpr int ST = decltype(stc)::value;
        constexpr int SOFF = ST * STAGE_FLOATS * 4, OOFF = (ST & 1) * OB_FLOATS * 4, POFF = ((ST & 1) ^ 1) * OB_FLOATS * 4;
        tma::mbar_wait(full0 + 8u * ST, (uint32_t)(i / NS) & 1u);
        if (NEED_GIN && i > 0) {
            // out-box of channel i - 1 -> grad_input plane (row-contiguous 16-byte reductions), re-zeroed
            float* gp = gin + (size_t)(plane0 + i - 1) * plane;
#pragma unroll
            for (int u = 0; u < QPT; ++u)
                if (quad_s[u] != 0xffffffffu) {
                    red_add4(gp + quad_g[u], lds4_i<POFF>(quad_s[u]));
                    sts4_i<POFF>(quad_s[u], make_float4(0.0f, 0.0f, 0.0f, 0.0f));
                }
        }
        if (gflow) {
            tma::static_for<PPT>([&](auto kc) {
                constexpr int k = decltype(kc)::value;
                const float g = lds_i<SOFF + 4 * ((k >> 1) * TW + (k & 1) * 32)>(a_g0);
                const float v_nw = lds_i<SOFF>(a_n[k]), v_ne = lds_i<SOFF>(a_n[k] + dx4[k]);
                const float v_sw = lds_i<SOFF>(a_s[k]), v_se = lds_i<SOFF>(a_s[k] + dx4[k]);
                // a tap outside the image re-reads its in-image neighbour with a zero 1-D weight
                const float tx = fmaf(w1[k][3], v_se - v_sw, w1[k][2] * (v_ne - v_nw));
                const float ty = fmaf(w1[k][1], v_se - v_ne, w1[k][0] * (v_sw - v_nw));
                gix[k] = fmaf(tx, g, gix[k]);
                giy[k] = fmaf(ty, g, giy[k]);
            });
        }
        if (NEED_GIN) {
            float acc = 0.0f;
#pragma unroll
            for (int j0 = 0; j0 < KMAX; j0 += GB) {  // gathers of a batch issued back to back, then the run sums
                float gv[GB];
#pragma unroll
                for (int u = 0; u < GB; ++u) gv[u] = lds_i<SOFF>(pa[j0 + u]);
#pragma unroll
                for (int u = 0; u < GB; ++u) {
                    acc = fmaf(pw[j0 + u], gv[u], acc);
                    const bool ends = (lastmask >> (j0 + u)) & 1u;
                    if (ends) sts_i<OOFF>(pd[j0 + u], acc);  // (predicated store, no branch)
                    acc = ends ? 0.0f : acc;
                }
            }
            // pieces of runs shared with the neighbouring shares (same thread wrote the slots).
            // (Completing a run cut once inside the warp by shuffle + plain store instead was
            // measured: 964 vs 943 us, the extra registers cost more than the atomics.)
            if (lastmask & (1u << 16)) red_shared(head_dst + OOFF, lds_i<OOFF>(ob_head));
            if (lastmask & (1u << 17)) red_shared(tail_dst + OOFF, lds_i<OOFF>(ob_head + 4u * THREADS));
        }
        __syncthreads();  // stage ST consumed by every thread; out-box i complete, out-box i - 1 added
        if (tid == 0 && i + NS < nch) issue_loads(i + NS);
    };
    for (int i0 = 0; i0 < nch; i0 += NS) {
        tma::static_for<NS>([&](auto stc) {
            if (i0 + decltype(stc)::value < nch) channel(stc, i0 + decltype(stc)::value);
        });
    }
    if (NEED_GIN) {
        float* gp = gin + (size_t)(plane0 + nch - 1) * plane;
        const uint32_t off = (uint32_t)((nch - 1) & 1) * (uint32_t)(OB_FLOATS * 4);
#pragma unroll
        for (int u = 0; u < QPT; ++u)
            if (quad_s[u] != 0xffffffffu) red_add4(gp + quad_g[u], lds4(quad_s[u] + off));
    }
    if (gflow) {
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            if (!valid[k]) continue;
            const int x = tx0 + (k & 1) * 32 + lane, y = ty0 + warp * 2 + (k >> 1);
            const size_t pix = (size_t)y * p.W + x;
            const float* fl = flow + (size_t)b * 2 * plane + pix;  // (recomputed: fewer live registers in the loop)
            const BwdCoord c2 = bwd_coord(__ldg(lin_x + x), __ldg(lin_y + y), __ldg(fl), __ldg(fl + plane), p);
            store_gflow(gflow, p, b, pix, c2, gix[k], giy[k], acc_gflow);
        }
    }
}

}  // namespace dsvc

using namespace dsvc;

static bool encode_xy_plane(CUtensorMap* tm, const float* base, const WarpParams& p, int box_w, int box_h) {
    auto encode = tensor_map_encoder();
    if (!encode) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B * p.C};
    const cuuint64_t gstride[2] = {(cuuint64_t)p.W * 4, (cuuint64_t)p.H * p.W * 4};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// returns -1 when the shape is not eligible (the caller uses the per-pixel kernel).
// grad_input must be zero on entry (it is accumulated into); grad_flow may be uninitialised.
int dsvc_warp_bwd_staged_launch(const float* gout, const float* input, const float* flow, float* gin,
                                float* gflow, const float* lin_x, const float* lin_y, const WarpParams& p,
                                bool force, cudaStream_t st, const int* mode) {
    if (p.W % 4 != 0 || !aligned16(input) || !aligned16(gout) || (gin && !aligned16(gin))) return -1;
    if (!force && (p.C < 8 || p.W < 64 || p.H < 16)) return -1;
    if ((long long)p.B * p.C > (1ll << 30)) return -1;
    CUtensorMap tm_in, tm_gout;
    if (!encode_xy_plane(&tm_in, input, p, bwd::BW, bwd::ROWCHUNK)) return -1;
    if (!encode_xy_plane(&tm_gout, gout, p, bwd::TW, bwd::TH)) return -1;
    static unsigned long long attr_set = 0;
    static int sms_of[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
    if (!((attr_set >> dev) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(warp_bwd_staged_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd::SMEM_BYTES);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(warp_bwd_staged_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bwd::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = DSVC_NUM_SMS;
        sms_of[dev] = n;
        attr_set |= 1ull << dev;
    }
    const int tiles_x = (p.W + bwd::TW - 1) / bwd::TW, tiles_y = (p.H + bwd::TH - 1) / bwd::TH;
    const long long ntiles = (long long)tiles_x * tiles_y * p.B;
    if (ntiles > (1ll << 24)) return -1;
    // fewer tiles than resident CTAs: cut the channels in ranges.  Not beyond that: the per-tile
    // set-up (CSR build, ~17 us) is repeated per range -- measured at 8x64x256x256 (512 tiles):
    // 1 range 280 us, 2 ranges 314 us, 4 ranges 337 us although 512 tiles are only 1.7 waves
    const int slots = 2 * sms_of[dev];
    int csplit = 1;
#ifdef DSVC_TUNE
    static int env_split = -1;
    if (env_split < 0) { const char* e = getenv("DSVC_BWD_CSPLIT"); env_split = e ? atoi(e) : 0; }
    if (env_split > 0) csplit = env_split;
    else
#endif
    while (ntiles * csplit < (long long)slots && p.C / (csplit * 2) >= 16) csplit *= 2;
    const int cper = (p.C + csplit - 1) / csplit;
    if (gflow && csplit > 1) {
        const cudaError_t e = cudaMemsetAsync(gflow, 0, (size_t)p.B * 2 * p.H * p.W * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
    }
    static int perm_mul = 1;  // lane -> share permutation multiplier (odd values measured: no effect)
#ifdef DSVC_TUNE
    static bool perm_read = false;
    if (!perm_read) { const char* e = getenv("DSVC_BWD_PERM"); perm_mul = e ? (atoi(e) | 1) : 1; perm_read = true; }
#endif
    const unsigned grid = (unsigned)(ntiles * csplit);
    if (gin)
        warp_bwd_staged_kernel<true><<<grid, bwd::THREADS, bwd::SMEM_BYTES, st>>>(
            tm_in, tm_gout, gout, input, flow, gin, gflow, lin_x, lin_y, p, tiles_x, tiles_y, csplit, cper, perm_mul, mode);
    else
        warp_bwd_staged_kernel<false><<<grid, bwd::THREADS, bwd::SMEM_BYTES, st>>>(
            tm_in, tm_gout, gout, input, flow, gin, gflow, lin_x, lin_y, p, tiles_x, tiles_y, csplit, cper, perm_mul, mode);
    return (int)cudaGetLastError();
}
