// Backward of the bilinear warp in CELL order (NCHW fp32, sm_100a): both gradients regular, no
// atomics on the main path, no zero-fill of grad_input.
//
// Replaces the autograd of /root/reference/modules.py:25-62 (ATen grid_sampler_2d_backward via
// modules.py:58-62, Learner.py:1343) for the wide (64-ch) feature warp.
//
// An output pixel p with clamped source coordinate (ix, iy) belongs to the CELL (X, Y) =
// (floor ix, floor iy) of the source image.  Seen from the cells, both gradients are regular:
//   * grad_input[Y][X] is the sum over the four cells (X-1..X, Y-1..Y) of their pixels'
//     grad_out times one of the pixel's four bilinear weights -- a thread that walks down a
//     column of cells gets the left neighbour's share by one shuffle and the upper row's share
//     from its own registers, and writes every element of grad_input exactly once (plain,
//     row-contiguous stores);
//   * the four input taps a pixel needs for grad_flow are input[Y..Y+1][X..X+1] of ITS CELL:
//     row-contiguous loads by cell coordinate, shared with the lane to the left by a shuffle.
// The only irregular access left is one load of grad_out per pixel and channel (pixels of
// adjacent cells are adjacent up to the flow's local distortion: near-coalesced, served by L1).
//
// Launches of one call (dsvc_warp_bwd_cell_launch), after two memsets (tables 0xFF, counters 0):
//   1. cell_build_kernel, one thread per output pixel: coordinates as in the forward (bit-exact
//      op order, warp_bwd_common.cuh), then the pixel claims layer 0 or 1 of its cell's table
//      entry by atomicCAS (pixel index + clamp flags, weights beside it).  A third, fourth ...
//      pixel of a cell is an EXTRA: it is appended to the bucket (<= 256 entries) of every
//      62 x 4R region of grad_input (one CTA of the next launch) that one of its taps reaches;
//      with a full bucket it goes to the launch-wide overflow list with the mask of the regions
//      that refused it.
//   2. the gradients, one warp per 31 x R block of grad_input (+1 halo column / row of cells),
//      lane = cell column, two table layers in registers, per channel 3R+2 shuffles and R row stores:
//      warp_bwd_cell_staged_kernel (default; section 2b) reads its 2(R+1) grad_out values and R+1
//      input values per channel with LDS from a TMA-fed ring; warp_bwd_cell_kernel (rows that are
//      not 16-byte multiples) with predicated global loads.  After a CTA's warps have stored their
//      blocks, its threads split (bucket entry, channel chunk) and add the extras' taps inside
//      the region with RED.ADD (and compute their grad_flow).
//   3. cell_fixup_kernel: the overflow list (empty for SpyNet-like flows), one warp per entry,
//      per-pixel scatter of exactly the taps nobody else took.
// Any flow is handled; the speed degrades towards the per-pixel kernel as cells fill up.  Like
// ATen's kernel the result is not bit-reproducible from run to run: which of a cell's pixels
// gets layer 0 is decided by the atomicCAS race, which permutes fp32 sums of two to four terms.
// Measured variants and what bounds the kernel: DESIGN.md 4.3.0, profiles/r02_bwd_cell_timing.txt.
#include "warp_bwd_common.cuh"
#include "tma_utils.cuh"
#include <cstdlib>

namespace dsvc {
namespace bcell {

constexpr int DC = 31;  // columns of grad_input per warp (lane 0 is the halo column of cells)
#ifndef DSVC_CELL_CTA_BY
#define DSVC_CELL_CTA_BY 4
#endif
#ifndef DSVC_CELL_STAGED_MINB
#define DSVC_CELL_STAGED_MINB 3  // CTAs per SM of the staged variant (R = 2)
#endif
constexpr int CTA_BX = 2, CTA_BY = DSVC_CELL_CTA_BY;  // a CTA's warps cover a region of 62 x CTA_BY R elements
constexpr int WARPS = CTA_BX * CTA_BY, THREADS = WARPS * 32;
constexpr int RX = CTA_BX * DC;
constexpr int EC = THREADS;            // bucket entries per region: one thread each
constexpr int D = 4;                   // channels in flight per thread (cp.async ring variant)
__host__ __device__ constexpr int slots_of(int R) { return 2 * (R + 1) + (R + 1); }  // grad_out of 2 layers x (R+1) rows, R+1 input rows
__host__ __device__ constexpr int ring_bytes_of(int R) { return D * slots_of(R) * THREADS * 4; }
constexpr int EMPTY = -1;
constexpr unsigned PIX_MASK = 0x3fffffffu, CLAMP_X = 0x40000000u, CLAMP_Y = 0x80000000u, NOPIX = 0xffffffffu;

// predicated accesses at base[off] (the address is formed inside: no 64-bit temporaries stay live;
// a false predicate never dereferences it)
__device__ __forceinline__ float ldg_valid(const float* base, unsigned off) {  // 0 where off == NOPIX
    float v;
    asm("{\n .reg .pred p;\n .reg .u64 a;\n setp.ne.u32 p, %2, 0xffffffff;\n mad.wide.u32 a, %2, 4, %1;\n mov.f32 %0, 0f00000000;\n"
        " @p ld.global.nc.f32 %0, [a];\n}"
        : "=f"(v)
        : "l"(base), "r"(off));
    return v;
}
__device__ __forceinline__ float ldg_off_if(const float* base, unsigned off, bool pred) {
    float v;
    asm("{\n .reg .pred p;\n .reg .u64 a;\n setp.ne.u32 p, %3, 0;\n mad.wide.u32 a, %2, 4, %1;\n mov.f32 %0, 0f00000000;\n"
        " @p ld.global.nc.f32 %0, [a];\n}"
        : "=f"(v)
        : "l"(base), "r"(off), "r"((unsigned)pred));
    return v;
}
__device__ __forceinline__ void cp_async4_if(unsigned dst_shared, const float* base, unsigned off, bool pred) {
    asm volatile(
        "{\n .reg .pred p;\n .reg .u64 a;\n setp.ne.u32 p, %3, 0;\n mad.wide.u32 a, %2, 4, %1;\n"
        " @p cp.async.ca.shared.global [%0], [a], 4;\n}" ::"r"(dst_shared),
        "l"(base), "r"(off), "r"((unsigned)pred)
        : "memory");
}
__device__ __forceinline__ void cp_async4_valid(unsigned dst_shared, const float* base, unsigned off) {  // off != NOPIX
    asm volatile(
        "{\n .reg .pred p;\n .reg .u64 a;\n setp.ne.u32 p, %2, 0xffffffff;\n mad.wide.u32 a, %2, 4, %1;\n"
        " @p cp.async.ca.shared.global [%0], [a], 4;\n}" ::"r"(dst_shared),
        "l"(base), "r"(off)
        : "memory");
}
__device__ __forceinline__ void st_off_if(float* base, unsigned off, float v, bool pred) {
    asm volatile(
        "{\n .reg .pred p;\n .reg .u64 a;\n setp.ne.u32 p, %3, 0;\n mad.wide.u32 a, %1, 4, %0;\n"
        " @p st.global.L1::no_allocate.f32 [a], %2;\n}" ::"l"(base),
        "r"(off), "f"(v), "r"((unsigned)pred)
        : "memory");
}
__device__ __forceinline__ void red_if(float* ptr, float v, bool pred) {
    asm volatile("{\n .reg .pred p;\n setp.ne.u32 p, %2, 0;\n @p red.global.add.f32 [%0], %1;\n}" ::"l"(ptr), "f"(v),
                 "r"((unsigned)pred)
                 : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
    asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
}

struct Layout {
    int nbx, nby;                     // 31 x R blocks (one warp each)
    int nrx, nry;                     // 62 x 4R regions (one CTA each)
    size_t counts_off, counts_bytes;  // int32: [0] overflow entries, [4 + region] bucket fill, both minus one (0xFF-filled per call with the tables behind them)
    size_t tab_off, tab_bytes;        // int32 [2][B][H*W] pixel word per cell and layer (0xFF-filled per call)
    size_t wts_off;                   // float2 [2][B][H*W] (wx1, wy1) of that pixel
    size_t bent_off;                  // int4 [regions][EC]
    size_t ovf_off;                   // int2 [B*H*W]
    size_t total;
};

static inline size_t up256(size_t v) { return (v + 255) & ~(size_t)255; }

static Layout make_layout(int B, int H, int W, int R) {
    Layout L;
    L.nbx = (W + DC - 1) / DC;
    L.nby = (H + R - 1) / R;
    L.nrx = (L.nbx + CTA_BX - 1) / CTA_BX;
    L.nry = (L.nby + CTA_BY - 1) / CTA_BY;
    const size_t px = (size_t)B * H * W, regions = (size_t)B * L.nrx * L.nry;
    size_t o = 0;
    L.counts_off = o;
    L.counts_bytes = up256((4 + regions) * sizeof(int));
    o += L.counts_bytes;
    L.tab_off = o;
    L.tab_bytes = up256(2 * px * sizeof(int));
    o += L.tab_bytes;
    L.wts_off = o;
    o += up256(2 * px * sizeof(float2));
    L.bent_off = o;
    o += up256(regions * EC * sizeof(int4));
    L.ovf_off = o;
    o += up256(px * sizeof(int2));
    L.total = o;
    return L;
}

// ---------------------------------------------------------------------------------- 1. tables
__global__ void __launch_bounds__(256)
cell_build_kernel(const float* __restrict__ flow, const float* __restrict__ lin_x, const float* __restrict__ lin_y,
                  WarpParams p, int* __restrict__ tab, float2* __restrict__ wts, int* __restrict__ counts,
                  int4* __restrict__ bent, int2* __restrict__ ovf, int nrx, int nry, int RY) {
    const int x = blockIdx.x * blockDim.x + threadIdx.x, y = blockIdx.y, b = blockIdx.z;
    if (x >= p.W) return;
    const size_t plane = (size_t)p.H * p.W, lay = (size_t)p.B * plane;
    const int pix = y * p.W + x;
    const float* fl = flow + (size_t)b * 2 * plane + pix;
    const BwdCoord c = bwd_coord(__ldg(lin_x + x), __ldg(lin_y + y), __ldg(fl), __ldg(fl + plane), p);
    const int X = c.t.x0, Y = c.t.y0;
    const unsigned word = (unsigned)pix | (c.gx_mult == 0.0f ? CLAMP_X : 0u) | (c.gy_mult == 0.0f ? CLAMP_Y : 0u);
    const size_t cell = (size_t)b * plane + (size_t)Y * p.W + X;
    int layer = 2;
    if (atomicCAS(tab + cell, EMPTY, (int)word) == EMPTY) layer = 0;
    else if (atomicCAS(tab + lay + cell, EMPTY, (int)word) == EMPTY) layer = 1;
    if (layer < 2) {
        wts[(size_t)layer * lay + cell] = make_float2(c.wx1, c.wy1);
        return;
    }
    // an extra: one bucket entry in every region that owns one of its four destinations
    const int rx = X / RX, ry = Y / RY;
    const bool ex = X + 1 < p.W && (X + 1) / RX != rx, ey = Y + 1 < p.H && (Y + 1) / RY != ry;
    int fail = 0;
    const int4 ent = make_int4((int)word, __float_as_int(c.wx1), __float_as_int(c.wy1), X | (Y << 16));
    auto put = [&](int rxx, int ryy, int bit) {
        const int reg = (b * nry + ryy) * nrx + rxx;
        const int k = atomicAdd(counts + 4 + reg, 1) + 1;  // counters start at -1
        if (k < EC) bent[(size_t)reg * EC + k] = ent;
        else fail |= bit;
    };
    put(rx, ry, 1);
    if (ex) put(rx + 1, ry, 2);
    if (ey) put(rx, ry + 1, 4);
    if (ex && ey) put(rx + 1, ry + 1, 8);
    if (fail) {
        const int k = atomicAdd(counts, 1) + 1;
        ovf[k] = make_int2((int)((size_t)b * plane + pix), fail);
    }
}

// ---------------------------------------------------------------------------------- 2. the gradients

// One channel of one warp's block: the four shares of every cell, combined across lanes (west
// neighbour) and rows (registers), stored row by row; grad_flow sums of the block's own cells.
// per cell: S0 = sum g, Sx = sum g wx1, Sy = sum g wy1, Sxy = sum g wx1 wy1 over its pixels;
// south-east share = Sxy, south-west = Sy - Sxy, north-east = Sx - Sxy, north-west = the rest
template <int R, bool NEED_GFLOW>
__device__ __forceinline__ void cell_channel(const float (&g)[R + 1][2], const float (&inR)[R + 1], const float (&wx)[R + 1][2],
                                             const float (&wy)[R + 1][2], float (&gx)[R][2], float (&gy)[R][2], float* ob,
                                             unsigned o_out, bool st_col, int rows_out, int W) {
    float bprev;
    {   // halo row of cells: only its south taps reach this block
        const float a0 = g[0][0] * wx[0][0], a1 = g[0][1] * wx[0][1];
        const float sy = fmaf(g[0][1], wy[0][1], g[0][0] * wy[0][0]);
        const float sxy = fmaf(a1, wy[0][1], a0 * wy[0][0]);
        bprev = (sy - sxy) + __shfl_up_sync(0xffffffffu, sxy, 1);
    }
    float inL0 = 0.0f, inR0 = 0.0f, dxt = 0.0f;
    if (NEED_GFLOW) {
        inR0 = inR[0];
        inL0 = __shfl_up_sync(0xffffffffu, inR0, 1);
        dxt = inR0 - inL0;
    }
#pragma unroll
    for (int lr = 1; lr <= R; ++lr) {
        const float a0 = g[lr][0] * wx[lr][0], a1 = g[lr][1] * wx[lr][1];
        const float s0 = g[lr][0] + g[lr][1], sx = a0 + a1;
        const float sy = fmaf(g[lr][1], wy[lr][1], g[lr][0] * wy[lr][0]);
        const float sxy = fmaf(a1, wy[lr][1], a0 * wy[lr][0]);
        const float bl = sy - sxy, tr = sx - sxy, tl = (s0 - sx) - bl;
        const float val = (tl + __shfl_up_sync(0xffffffffu, tr, 1)) + bprev;
        bprev = bl + __shfl_up_sync(0xffffffffu, sxy, 1);
        st_off_if(ob, o_out + (unsigned)((lr - 1) * W), val, st_col && lr - 1 < rows_out);
        if (NEED_GFLOW) {
            const float inR1 = inR[lr];
            const float inL1 = __shfl_up_sync(0xffffffffu, inR1, 1);
            const float dxb = inR1 - inL1, dyl = inL1 - inL0, dyr = inR1 - inR0;
            const float ddx = dxb - dxt, ddy = dyr - dyl;
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                gx[lr - 1][e] = fmaf(g[lr][e], fmaf(wy[lr][e], ddx, dxt), gx[lr - 1][e]);
                gy[lr - 1][e] = fmaf(g[lr][e], fmaf(wx[lr][e], ddy, dyl), gy[lr - 1][e]);
            }
            inR0 = inR1;
            inL0 = inL1;
            dxt = dxb;
        }
    }
}

// grad_flow = clip multiplier * channel sum / scale  (store_gflow, warp_bwd_common.cuh)
__device__ __forceinline__ void cell_put_gflow(float* __restrict__ gflow, const WarpParams& p, int b, bool accumulate,
                                               unsigned pix, unsigned cl, float sx_, float sy_) {
    const size_t plane = (size_t)p.H * p.W;
    float* gf = gflow + (size_t)b * 2 * plane;
    const float mx = (float)(p.W - 1) * 0.5f, my = (float)(p.H - 1) * 0.5f;
    const float ggx = __fmul_rn((cl & 1u) ? 0.0f : mx, sx_), ggy = __fmul_rn((cl & 2u) ? 0.0f : my, sy_);
    const float vx = p.flow_mode ? __fdiv_rn(ggx, p.sx) : __fmul_rn(ggx, p.inv_sx);
    const float vy = p.flow_mode ? __fdiv_rn(ggy, p.sy) : __fmul_rn(ggy, p.inv_sy);
    if (accumulate) {
        atomicAdd(gf + pix, vx);
        atomicAdd(gf + plane + pix, vy);
    } else {
        gf[pix] = vx;
        gf[plane + pix] = vy;
    }
}

// The region's extras, behind every warp's stores (call with all threads of the CTA): the CTA's
// threads split the bucket entries AND the channels (entry e, chunk k of the channel range), each
// adds its taps inside the region with RED.ADD; the chunks' grad_flow sums meet in shared memory.
template <int R, bool NEED_GFLOW>
__device__ __forceinline__ void cell_region_extras(const float* __restrict__ gout, const float* __restrict__ in,
                                                   float* __restrict__ gin, float* __restrict__ gflow,
                                                   const int* __restrict__ counts, const int4* __restrict__ bent,
                                                   const WarpParams& p, int b, int c0, int c1, bool accumulate) {
    const int W = p.W, H = p.H;
    const size_t plane = (size_t)H * W;
    const int region = (b * (int)gridDim.y + (int)blockIdx.y) * (int)gridDim.x + (int)blockIdx.x;
    const int n_ex = min(__ldg(counts + 4 + region) + 1, EC);
    if (n_ex == 0 || c0 >= c1) return;  // CTA-uniform
    __shared__ float s_gx[EC], s_gy[EC];
    for (int i = threadIdx.x; i < n_ex; i += blockDim.x) s_gx[i] = s_gy[i] = 0.0f;
    __syncthreads();
    const int nchunk = max(1, min((int)blockDim.x / n_ex, (c1 - c0 + 3) / 4));  // >= 4 channels per chunk
    const int per = (((c1 - c0) + nchunk - 1) / nchunk + 3) & ~3;
    const int ei = (int)threadIdx.x % n_ex, k = (int)threadIdx.x / n_ex;
    const int ca = c0 + k * per, cb = min(c1, ca + per);
    bool own = false;
    unsigned pix = 0, clampf = 0;
    if (k < nchunk && ca < cb) {
        const int4 e = __ldg(bent + (size_t)region * EC + ei);
        pix = (unsigned)e.x & PIX_MASK;
        clampf = (unsigned)e.x >> 30;
        const float ewx = __int_as_float(e.y), ewy = __int_as_float(e.z);
        const int cx = e.w & 0xffff, cy = (int)((unsigned)e.w >> 16);
        constexpr int RY = CTA_BY * R;
        const int rx = (int)blockIdx.x, ry = (int)blockIdx.y;
        const bool xe = cx + 1 < W, ys = cy + 1 < H;
        const bool cx0 = cx / RX == rx, cx1 = xe && (cx + 1) / RX == rx, cy0 = cy / RY == ry, cy1 = ys && (cy + 1) / RY == ry;
        const bool t0 = cx0 && cy0, t1 = cx1 && cy0, t2 = cx0 && cy1, t3 = cx1 && cy1;
        own = NEED_GFLOW && t0;  // the region of the cell's own element computes the pixel's grad_flow
        const unsigned ebase = (unsigned)(cy * W + cx);
        const float w_se = ewx * ewy, w_sw = ewy - w_se, w_ne = ewx - w_se, w_nw = (1.0f - ewx) - w_sw;
        float egx = 0.0f, egy = 0.0f;
        const float* gp = gout + ((size_t)b * p.C + ca) * plane + pix;
        const float* ip = in + ((size_t)b * p.C + ca) * plane + ebase;
        float* op = gin + ((size_t)b * p.C + ca) * plane + ebase;
        constexpr int U = 4;
        for (int c = ca; c < cb; c += U) {
            float ge[U], v00[U], v01[U], v10[U], v11[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool cok = c + u < cb;
                const size_t o = (size_t)u * plane;
                ge[u] = cok ? __ldg(gp + o) : 0.0f;
                if (own) {
                    v00[u] = cok ? __ldg(ip + o) : 0.0f;
                    v01[u] = (cok && xe) ? __ldg(ip + o + 1) : 0.0f;
                    v10[u] = (cok && ys) ? __ldg(ip + o + W) : 0.0f;
                    v11[u] = (cok && xe && ys) ? __ldg(ip + o + W + 1) : 0.0f;
                }
            }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const bool cok = c + u < cb;
                float* o = op + (size_t)u * plane;
                red_if(o, ge[u] * w_nw, cok && t0);
                red_if(o + 1, ge[u] * w_ne, cok && t1);
                red_if(o + W, ge[u] * w_sw, cok && t2);
                red_if(o + W + 1, ge[u] * w_se, cok && t3);
                if (own) {
                    const float dxt = v01[u] - v00[u], dxb = v11[u] - v10[u], dyl = v10[u] - v00[u], dyr = v11[u] - v01[u];
                    egx = fmaf(ge[u], fmaf(ewy, dxb - dxt, dxt), egx);
                    egy = fmaf(ge[u], fmaf(ewx, dyr - dyl, dyl), egy);
                }
            }
            gp += (size_t)U * plane;
            ip += (size_t)U * plane;
            op += (size_t)U * plane;
        }
        if (own) {
            atomicAdd(&s_gx[ei], egx);
            atomicAdd(&s_gy[ei], egy);
        }
    }
    if (!NEED_GFLOW) return;
    __syncthreads();
    if (own && k == 0) cell_put_gflow(gflow, p, b, accumulate, pix, clampf, s_gx[ei], s_gy[ei]);
}

template <int R, bool NEED_GFLOW, int MINB, bool RING>
__global__ void __launch_bounds__(THREADS, MINB)
warp_bwd_cell_kernel(const float* __restrict__ gout, const float* __restrict__ in, float* __restrict__ gin,
                     float* __restrict__ gflow, const int* __restrict__ tab, const float2* __restrict__ wts,
                     const int* __restrict__ counts, const int4* __restrict__ bent, WarpParams p, int nbx, int nby,
                     int csplit, int c_per) {
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int bxi = blockIdx.x * CTA_BX + (warp % CTA_BX), byi = blockIdx.y * CTA_BY + warp / CTA_BX;
    const int b = blockIdx.z / csplit, split = blockIdx.z - b * csplit;
    const int c0 = split * c_per, c1 = min(p.C, c0 + c_per);
    const int W = p.W, H = p.H;
    const size_t plane = (size_t)H * W, lay = (size_t)p.B * plane;
    if (bxi < nbx && byi < nby && c0 < c1) {  // warp-uniform
        const int Xb = bxi * DC, Yb = byi * R, X = Xb - 1 + lane;
        const bool colok = X >= 0 && X < W;
        // two layers of (pixel, weights) per cell of this lane's column, rows Yb-1 .. Yb+R-1
        unsigned off[R + 1][2];  // pixel index in the plane, NOPIX where the layer is empty
        float wx[R + 1][2], wy[R + 1][2];
#pragma unroll
        for (int lr = 0; lr <= R; ++lr) {
            const int Y = Yb - 1 + lr;
            const bool ok = colok && Y >= 0 && Y < H;
            const size_t cell = (size_t)b * plane + (ok ? (size_t)Y * W + X : 0);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                const int word = ok ? __ldg(tab + e * lay + cell) : EMPTY;
                const bool v = word != EMPTY;
                const float2 w = v ? __ldg(wts + e * lay + cell) : make_float2(0.0f, 0.0f);
                off[lr][e] = v ? ((unsigned)word & PIX_MASK) : NOPIX;
                wx[lr][e] = w.x;
                wy[lr][e] = w.y;
            }
        }
        float gx[R][2], gy[R][2];
#pragma unroll
        for (int r = 0; r < R; ++r) gx[r][0] = gx[r][1] = gy[r][0] = gy[r][1] = 0.0f;

        const bool in_col = X + 1 < W;           // the input column this lane loads: its cells' east taps
        const bool st_col = lane >= 1 && X < W;  // lane 0 only feeds lane 1
        const unsigned o_in = (unsigned)(Yb * W + X + 1), o_out = (unsigned)(Yb * W + X);
        const int rows_in = min(R + 1, H - Yb), rows_out = min(R, H - Yb);  // rows of this block inside the image

        // RING: every thread streams its own values of a channel through a private D-deep ring in shared
        // memory with cp.async: the loads of channel c+D-1 are in flight while channel c is computed, no
        // registers are held for them and no thread reads another's slot (so no barrier).  Slots of empty
        // layers are zeroed once and never written.  !RING: predicated loads at the top of the iteration.
        constexpr int SLOTS = slots_of(R);
        extern __shared__ float ring[];  // [D][SLOTS][THREADS]
        float* mine = ring + threadIdx.x;
        unsigned mine_s = 0;
        if (RING) {
#pragma unroll
            for (int i = 0; i < D * SLOTS; ++i) mine[i * THREADS] = 0.0f;
            mine_s = (unsigned)__cvta_generic_to_shared(mine);
        }
        const float* gb = gout + ((size_t)b * p.C + c0) * plane;  // channel being loaded / enqueued
        const float* ib = in + ((size_t)b * p.C + c0) * plane;
        float* ob = gin + ((size_t)b * p.C + c0) * plane;          // channel being computed
        auto enqueue = [&](int stage) {
            const unsigned dst = mine_s + (unsigned)(stage * SLOTS * THREADS * 4);
#pragma unroll
            for (int lr = 0; lr <= R; ++lr)
#pragma unroll
                for (int e = 0; e < 2; ++e) cp_async4_valid(dst + (lr * 2 + e) * THREADS * 4, gb, off[lr][e]);
            if (NEED_GFLOW) {
#pragma unroll
                for (int k = 0; k <= R; ++k)
                    cp_async4_if(dst + (2 * (R + 1) + k) * THREADS * 4, ib, o_in + (unsigned)(k * W), in_col && k < rows_in);
            }
            gb += plane;
            ib += plane;
        };
        int c_enq = c0;
        if (RING) {
#pragma unroll
            for (int d = 0; d < D - 1; ++d) {
                if (c_enq < c1) enqueue(d);
                cp_async_commit();
                ++c_enq;
            }
        }
        int st_rd = 0, st_wr = D - 1;

        for (int c = c0; c < c1; ++c) {
            float g[R + 1][2], inR[R + 1];
            if (RING) {
                if (c_enq < c1) enqueue(st_wr);
                cp_async_commit();
                ++c_enq;
                st_wr = st_wr == D - 1 ? 0 : st_wr + 1;
                cp_async_wait<D - 1>();
                const float* rs = mine + st_rd * SLOTS * THREADS;
                st_rd = st_rd == D - 1 ? 0 : st_rd + 1;
#pragma unroll
                for (int lr = 0; lr <= R; ++lr)
#pragma unroll
                    for (int e = 0; e < 2; ++e) g[lr][e] = rs[(lr * 2 + e) * THREADS];
                if (NEED_GFLOW) {
#pragma unroll
                    for (int k = 0; k <= R; ++k) inR[k] = rs[(2 * (R + 1) + k) * THREADS];
                }
            } else {
#pragma unroll
                for (int lr = 0; lr <= R; ++lr)
#pragma unroll
                    for (int e = 0; e < 2; ++e) g[lr][e] = ldg_valid(gb, off[lr][e]);
                if (NEED_GFLOW) {
#pragma unroll
                    for (int k = 0; k <= R; ++k) inR[k] = ldg_off_if(ib, o_in + (unsigned)(k * W), in_col && k < rows_in);
                }
                gb += plane;
                ib += plane;
            }

            cell_channel<R, NEED_GFLOW>(g, inR, wx, wy, gx, gy, ob, o_out, st_col, rows_out, W);
            ob += plane;
        }

        if (NEED_GFLOW && lane >= 1) {
#pragma unroll
            for (int lr = 1; lr <= R; ++lr) {
                const int Y = Yb - 1 + lr;
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (off[lr][e] != NOPIX) {  // the clamp flags sit beside the pixel index in the table
                        const unsigned word = (unsigned)__ldg(tab + e * lay + (size_t)b * plane + (size_t)Y * W + X);
                        cell_put_gflow(gflow, p, b, csplit > 1, off[lr][e], word >> 30, gx[lr - 1][e], gy[lr - 1][e]);
                    }
            }
        }
    }

    cell_region_extras<R, NEED_GFLOW>(gout, in, gin, gflow, counts, bent, p, b, c0, c1, csplit > 1);
}



// ---------------------------------------------------------------------------------- 2b. staged variant
// The same block arithmetic fed from shared memory.  The pixels of a region's cells lie in a
// bounding box of grad_out (the region moved by minus the local flow: 68 x 20 at the median, 88 x 32
// at most under the synthetic SpyNet-like flow); when that box fits a stage, a producer warp streams
// it, and the region's rows of the input, channel by channel into an NS-deep ring with TMA tensor
// loads (two box sizes of grad_out, one request per box, completing on an mbarrier), and the eight
// block warps read their values with LDS at fixed offsets: no per-load address arithmetic, no
// predicates (empty layers point at a zero word; TMA zero-fills outside the image), and the number
// of scattered global loads in flight -- what bounds the load variant -- no longer matters.  A
// region whose box does not fit runs the load variant's loop inside this kernel.
constexpr int S_WARPS = WARPS + 1, S_THREADS = S_WARPS * 32;  // 8 block warps and the producer
constexpr int NS = 4;                 // ring stages (channels in flight)
constexpr int GS_W = 72, GL_W = 88;   // small / large grad_out box widths (floats)
__host__ __device__ constexpr int gs_h(int R) { return CTA_BY * R + 12; }
__host__ __device__ constexpr int gl_h(int R) { return CTA_BY * R + 20; }
constexpr int IP = 72;                // input box width (63 columns + alignment slack)
__host__ __device__ constexpr int s_in_rows(int R) { return CTA_BY * R + 1; }
__host__ __device__ constexpr int s_g_bytes(int R) { return (GL_W * gl_h(R) * 4 + 127) & ~127; }  // the zero words sit behind it
__host__ __device__ constexpr int s_stage_bytes(int R) { return (s_g_bytes(R) + 128 + IP * s_in_rows(R) * 4 + 127) & ~127; }

// The compiler re-materialises shared-window addresses and index arithmetic inside the channel loop
// rather than spend a register on them (~30 of 155 instructions per channel); an empty asm makes the
// value opaque, so it stays where it is.
__device__ __forceinline__ void keep_in_register(unsigned& v) { asm volatile("" : "+r"(v)); }
__device__ __forceinline__ float lds_at(unsigned addr) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(addr));
    return v;
}

template <int R, bool NEED_GFLOW, int MINB>
__global__ void __launch_bounds__(S_THREADS, MINB)
warp_bwd_cell_staged_kernel(const __grid_constant__ CUtensorMap tm_gs, const __grid_constant__ CUtensorMap tm_gl,
                            const __grid_constant__ CUtensorMap tm_in, const float* __restrict__ gout,
                            const float* __restrict__ in, float* __restrict__ gin, float* __restrict__ gflow,
                            const int* __restrict__ tab, const float2* __restrict__ wts, const int* __restrict__ counts,
                            const int4* __restrict__ bent, WarpParams p, int nbx, int nby, int csplit, int c_per) {
    using namespace tma;
    constexpr int STAGE = s_stage_bytes(R), S_ZERO = s_g_bytes(R), S_IN = s_g_bytes(R) + 128;
    extern __shared__ __align__(128) unsigned char ring_raw[];
    __shared__ uint64_t full_bar[NS], empty_bar[NS];
    __shared__ int bbox[4];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const bool is_main = warp < WARPS;
    const int bxi = blockIdx.x * CTA_BX + (warp % CTA_BX), byi = blockIdx.y * CTA_BY + warp / CTA_BX;
    const int b = blockIdx.z / csplit, split = blockIdx.z - b * csplit;
    const int c0 = split * c_per, c1 = min(p.C, c0 + c_per);
    const int W = p.W, H = p.H;
    const size_t plane = (size_t)H * W, lay = (size_t)p.B * plane;
    const bool active = is_main && bxi < nbx && byi < nby && c0 < c1;  // warp-uniform
    // active block warps of this CTA (they release the stages)
    const int act_x = min(CTA_BX, nbx - (int)blockIdx.x * CTA_BX), act_y = min(CTA_BY, nby - (int)blockIdx.y * CTA_BY);
    const int n_act = c0 < c1 ? act_x * act_y : 0;
    const unsigned ring_s = smem_u32(ring_raw);

    if (threadIdx.x == 0) {
        for (int i = 0; i < NS; ++i) {
            mbar_init(&full_bar[i], 1);
            mbar_init(&empty_bar[i], (uint32_t)max(n_act, 1));
        }
        bbox[0] = 0x7fffffff; bbox[1] = 0x7fffffff; bbox[2] = -1; bbox[3] = -1;
        fence_barrier_init();
    }
    if (threadIdx.x < NS * 32) reinterpret_cast<float*>(ring_raw + (threadIdx.x >> 5) * STAGE + S_ZERO)[lane] = 0.0f;

    const int Xb = bxi * DC, Yb = byi * R, X = Xb - 1 + lane;
    unsigned off[R + 1][2], pxy[R + 1][2], clampb = 0;  // pixel index, its (x, y), clamp flags (2 bits per entry)
    float wx[R + 1][2], wy[R + 1][2];
    int mnx = 0x7fffffff, mny = 0x7fffffff, mxx = -1, mxy = -1;
    if (active) {
        const bool colok = X >= 0 && X < W;
#pragma unroll
        for (int lr = 0; lr <= R; ++lr) {
            const int Y = Yb - 1 + lr;
            const bool ok = colok && Y >= 0 && Y < H;
            const size_t cell = (size_t)b * plane + (ok ? (size_t)Y * W + X : 0);
#pragma unroll
            for (int e = 0; e < 2; ++e) {
                // both table reads are issued together (the weights of an empty layer are never used)
                const int word = ok ? __ldg(tab + e * lay + cell) : EMPTY;
                const float2 w = ok ? __ldg(wts + e * lay + cell) : make_float2(0.0f, 0.0f);
                const bool v = word != EMPTY;
                off[lr][e] = v ? ((unsigned)word & PIX_MASK) : NOPIX;
                wx[lr][e] = v ? w.x : 0.0f;
                wy[lr][e] = v ? w.y : 0.0f;
                pxy[lr][e] = 0;
                if (v) {
                    const int py = (int)(off[lr][e] / (unsigned)W), px = (int)off[lr][e] - py * W;
                    pxy[lr][e] = (unsigned)px | ((unsigned)py << 16);
                    clampb |= ((unsigned)word >> 30) << (lr * 4 + e * 2);
                    mnx = min(mnx, px); mxx = max(mxx, px); mny = min(mny, py); mxy = max(mxy, py);
                }
            }
        }
        mnx = __reduce_min_sync(0xffffffffu, mnx); mny = __reduce_min_sync(0xffffffffu, mny);
        mxx = __reduce_max_sync(0xffffffffu, mxx); mxy = __reduce_max_sync(0xffffffffu, mxy);
    }
    __syncthreads();  // barriers and the box are initialised
    if (active && lane == 0 && mxx >= 0) {
        atomicMin(&bbox[0], mnx); atomicMin(&bbox[1], mny); atomicMax(&bbox[2], mxx); atomicMax(&bbox[3], mxy);
    }
    __syncthreads();
    // the box of grad_out this region needs (columns from a 16-byte boundary); empty when no cell has a pixel
    const bool any = bbox[2] >= 0;
    const int bx0 = any ? (bbox[0] & ~3) : 0, by0 = any ? bbox[1] : 0;
    const int bw = any ? bbox[2] - bx0 + 1 : 0, bh = any ? bbox[3] - by0 + 1 : 0;
    const bool small = bw <= GS_W && bh <= gs_h(R);
    const bool staged = bw <= GL_W && bh <= gl_h(R);  // CTA-uniform
    const int pitch = small ? GS_W : GL_W;
    // the input rows of the region: columns from the 16-byte boundary at or below its first column
    const int Xr = (int)blockIdx.x * RX, Yr = (int)blockIdx.y * CTA_BY * R;
    const int ix0 = Xr & ~3;

    float gx[R][2], gy[R][2];
#pragma unroll
    for (int r = 0; r < R; ++r) gx[r][0] = gx[r][1] = gy[r][0] = gy[r][1] = 0.0f;
    const bool st_col = lane >= 1 && X < W;
    const unsigned o_out = (unsigned)(Yb * W + X);
    const int rows_out = min(R, H - Yb);

    if (staged) {
        if (warp == WARPS) {
            // ---- producer warp: two tensor loads per channel (OOB elements arrive as zeros)
            if (lane == 0 && n_act > 0) {
                const unsigned total = (unsigned)((small ? GS_W * gs_h(R) : GL_W * gl_h(R)) * 4 + (NEED_GFLOW ? IP * s_in_rows(R) * 4 : 0));
                const CUtensorMap* tg = small ? &tm_gs : &tm_gl;
                for (int i = 0; c0 + i < c1; ++i) {
                    const int s = i % NS;
                    if (i >= NS) mbar_wait(smem_u32(&empty_bar[s]), (unsigned)((i / NS - 1) & 1));
                    const unsigned bar = smem_u32(&full_bar[s]), dst = ring_s + (unsigned)(s * STAGE);
                    mbar_arrive_expect_tx(bar, total);
                    load_3d(dst, tg, bx0, by0, b * p.C + c0 + i, bar);
                    if (NEED_GFLOW) load_3d(dst + (unsigned)S_IN, &tm_in, ix0, Yr, b * p.C + c0 + i, bar);
                }
            }
        } else if (active) {
            // ---- block warps: shared-window addresses of the two layers' pixels and of the input column
            unsigned sa[R + 1][2];
#pragma unroll
            for (int lr = 0; lr <= R; ++lr)
#pragma unroll
                for (int e = 0; e < 2; ++e) {
                    unsigned a = S_ZERO;
                    if (off[lr][e] != NOPIX) {
                        const int py = (int)(pxy[lr][e] >> 16), px = (int)(pxy[lr][e] & 0xffffu);
                        a = (unsigned)(((py - by0) * pitch + (px - bx0)) * 4);
                    }
                    sa[lr][e] = ring_s + a;
                    keep_in_register(sa[lr][e]);
                }
            unsigned ia = ring_s + (unsigned)(S_IN + ((Yb - Yr) * IP + (X + 1 - ix0)) * 4);
            unsigned full0 = smem_u32(&full_bar[0]), empty0 = smem_u32(&empty_bar[0]), o_out_r = o_out;
            unsigned st_rows = st_col ? (unsigned)max(rows_out, 0) : 0u;  // rows this lane stores
            keep_in_register(ia);
            keep_in_register(full0);
            keep_in_register(empty0);
            keep_in_register(o_out_r);
            keep_in_register(st_rows);
            float* ob = gin + ((size_t)b * p.C + c0) * plane;
            for (int i0 = 0; c0 + i0 < c1; i0 += NS) {
                static_for<NS>([&](auto sc) {
                    constexpr int s = decltype(sc)::value;
                    if (c0 + i0 + s < c1) {
                        mbar_wait(full0 + 8 * s, (unsigned)((i0 / NS) & 1));
                        float g[R + 1][2], inR[R + 1];
#pragma unroll
                        for (int lr = 0; lr <= R; ++lr)
#pragma unroll
                            for (int e = 0; e < 2; ++e) g[lr][e] = lds_at(sa[lr][e] + s * STAGE);
                        if (NEED_GFLOW) {
#pragma unroll
                            for (int k = 0; k <= R; ++k) inR[k] = lds_at(ia + s * STAGE + k * IP * 4);
                        }
                        cell_channel<R, NEED_GFLOW>(g, inR, wx, wy, gx, gy, ob, o_out_r, true, (int)st_rows, W);
                        __syncwarp();
                        if (lane == 0) mbar_arrive(empty0 + 8 * s);
                        ob += plane;
                    }
                });
            }
            if (NEED_GFLOW && lane >= 1) {
                // the pixel index back from its shared-window address (the table words are not kept live)
#pragma unroll
                for (int lr = 1; lr <= R; ++lr)
#pragma unroll
                    for (int e = 0; e < 2; ++e) {
                        const unsigned a = (sa[lr][e] - ring_s) >> 2;
                        if (a != (unsigned)(S_ZERO >> 2)) {
                            const unsigned row = a / (unsigned)pitch, col = a - row * (unsigned)pitch;
                            cell_put_gflow(gflow, p, b, csplit > 1, (unsigned)((by0 + (int)row) * W + bx0 + (int)col),
                                           (clampb >> (lr * 4 + e * 2)) & 3u, gx[lr - 1][e], gy[lr - 1][e]);
                        }
                    }
            }
        }
    } else if (active) {
        // ---- the box does not fit: predicated loads (the load variant's loop)
        const bool in_col = X + 1 < W;
        const unsigned o_in = (unsigned)(Yb * W + X + 1);
        const int rows_in = min(R + 1, H - Yb);
        const float* gb = gout + ((size_t)b * p.C + c0) * plane;
        const float* ib = in + ((size_t)b * p.C + c0) * plane;
        float* ob = gin + ((size_t)b * p.C + c0) * plane;
        for (int c = c0; c < c1; ++c) {
            float g[R + 1][2], inR[R + 1];
#pragma unroll
            for (int lr = 0; lr <= R; ++lr)
#pragma unroll
                for (int e = 0; e < 2; ++e) g[lr][e] = ldg_valid(gb, off[lr][e]);
            if (NEED_GFLOW) {
#pragma unroll
                for (int k = 0; k <= R; ++k) inR[k] = ldg_off_if(ib, o_in + (unsigned)(k * W), in_col && k < rows_in);
            }
            cell_channel<R, NEED_GFLOW>(g, inR, wx, wy, gx, gy, ob, o_out, st_col, rows_out, W);
            gb += plane;
            ib += plane;
            ob += plane;
        }
        if (NEED_GFLOW && lane >= 1) {
#pragma unroll
            for (int lr = 1; lr <= R; ++lr) {
#pragma unroll
                for (int e = 0; e < 2; ++e)
                    if (off[lr][e] != NOPIX)
                        cell_put_gflow(gflow, p, b, csplit > 1, off[lr][e], (clampb >> (lr * 4 + e * 2)) & 3u, gx[lr - 1][e], gy[lr - 1][e]);
            }
        }
    }
    cell_region_extras<R, NEED_GFLOW>(gout, in, gin, gflow, counts, bent, p, b, c0, c1, csplit > 1);
}

// ---------------------------------------------------------------------------------- 3. overflow
template <bool NEED_GFLOW>
__global__ void __launch_bounds__(256)
cell_fixup_kernel(const float* __restrict__ gout, const float* __restrict__ in, const float* __restrict__ flow,
                  float* __restrict__ gin, float* __restrict__ gflow, const float* __restrict__ lin_x,
                  const float* __restrict__ lin_y, WarpParams p, const int* __restrict__ counts,
                  const int2* __restrict__ ovf, int RY) {
    const int n = __ldg(counts) + 1;
    const int lane = threadIdx.x & 31;
    const int nwarps = gridDim.x * (blockDim.x >> 5);
    const size_t plane = (size_t)p.H * p.W;
    for (int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5); i < n; i += nwarps) {
        const int2 e = __ldg(ovf + i);
        const int b = (int)((size_t)e.x / plane), pix = (int)((size_t)e.x - (size_t)b * plane);
        const int y = pix / p.W, x = pix - y * p.W;
        const float* fl = flow + (size_t)b * 2 * plane + pix;
        const BwdCoord bc = bwd_coord(__ldg(lin_x + x), __ldg(lin_y + y), __ldg(fl), __ldg(fl + plane), p);
        const int X = bc.t.x0, Y = bc.t.y0;
        const int rx = X / RX, ry = Y / RY;
        bool on[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int dx = t & 1, dy = t >> 1;
            const bool inimg = X + dx < p.W && Y + dy < p.H;
            const int bit = (((X + dx) / RX != rx) ? 1 : 0) + (((Y + dy) / RY != ry) ? 2 : 0);
            on[t] = inimg && ((e.y >> bit) & 1);
        }
        const bool own = e.y & 1;
        const int o_nw = Y * p.W + X;
        const bool xe = bc.t.x1ok, ys = bc.t.y1ok;
        float gix = 0.0f, giy = 0.0f;
        for (int c = lane; c < p.C; c += 32) {
            const size_t ch = ((size_t)b * p.C + c) * plane;
            const float g = __ldg(gout + ch + pix);
            float* gi = gin + ch + o_nw;
            if (on[0]) atomicAdd(gi, __fmul_rn(bc.t.nw, g));
            if (on[1]) atomicAdd(gi + 1, __fmul_rn(bc.t.ne, g));
            if (on[2]) atomicAdd(gi + p.W, __fmul_rn(bc.t.sw, g));
            if (on[3]) atomicAdd(gi + p.W + 1, __fmul_rn(bc.t.se, g));
            if (NEED_GFLOW && own) {
                const float* ip = in + ch + o_nw;
                const float v00 = __ldg(ip), v01 = xe ? __ldg(ip + 1) : 0.0f;
                const float v10 = ys ? __ldg(ip + p.W) : 0.0f, v11 = (xe && ys) ? __ldg(ip + p.W + 1) : 0.0f;
                gix += g * (bc.wy0 * (v01 - v00) + bc.wy1 * (v11 - v10));
                giy += g * (bc.wx0 * (v10 - v00) + bc.wx1 * (v11 - v01));
            }
        }
        if (NEED_GFLOW && own) {
            gix = warp_sum(gix);
            giy = warp_sum(giy);
            if (lane == 0) store_gflow(gflow, p, b, (size_t)pix, bc, gix, giy, false);
        }
    }
}

}  // namespace bcell
}  // namespace dsvc

using namespace dsvc;

#ifdef DSVC_TUNE
static int cell_knob(const char* name, int dflt) {
    const char* v = getenv(name);
    return v ? atoi(v) : dflt;
}
#endif
// Rows per warp.  Measured on B200 (staged variant, both gradients, 1 x 64 x 1088 x 1920 / 8 x 64 x 256 x 256):
// R = 2 at 3 CTAs per SM 650 / 220 us, R = 3 at 2 CTAs per SM 670 / 232 us, R = 4 730 / 243 us.
static int cell_rows(const WarpParams&) {
#ifdef DSVC_TUNE
    static const int r = cell_knob("DSVC_CELL_R", 0);
    if (r >= 2 && r <= 4) return r;
#endif
    return 2;
}

size_t dsvc_warp_bwd_cell_workspace(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    size_t m = 0;
    for (int r = 2; r <= 4; ++r) m = max(m, bcell::make_layout(B, H, W, r).total);
    return m;
}

template <int R, bool GF, int MB, bool RING>
static cudaError_t cell_launch_one(dim3 grid, cudaStream_t st, const float* gout, const float* input, float* gin, float* gflow,
                                   const int* tab, const float2* wts, const int* counts, const int4* bent, const WarpParams& p,
                                   int nbx, int nby, int csplit, int c_per) {
    using namespace bcell;
    constexpr int smem = RING ? ring_bytes_of(R) : 0;
    static bool attr_set[64] = {};
    int dev = 0;
    if (smem > 32 * 1024 && cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaFuncSetAttribute(warp_bwd_cell_kernel<R, GF, MB, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_set[dev] = true;
    }
    warp_bwd_cell_kernel<R, GF, MB, RING><<<grid, THREADS, smem, st>>>(gout, input, gin, gflow, tab, wts, counts, bent, p, nbx, nby,
                                                                     csplit, c_per);
    return cudaGetLastError();
}

static bool cell_encode_map(CUtensorMap* tm, const float* base, const WarpParams& p, int box_w, int box_h) {
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

// cudaErrorNotSupported when the tensor maps cannot be encoded (the caller launches the load variant)
template <int R, bool GF, int MB>
static cudaError_t cell_launch_staged(dim3 grid, cudaStream_t st, const float* gout, const float* input, float* gin, float* gflow,
                                      const int* tab, const float2* wts, const int* counts, const int4* bent, const WarpParams& p,
                                      int nbx, int nby, int csplit, int c_per) {
    using namespace bcell;
    CUtensorMap tm_gs, tm_gl, tm_in;
    if (!cell_encode_map(&tm_gs, gout, p, GS_W, gs_h(R)) || !cell_encode_map(&tm_gl, gout, p, GL_W, gl_h(R)) ||
        !cell_encode_map(&tm_in, input, p, IP, s_in_rows(R)))
        return cudaErrorNotSupported;
    constexpr int smem = NS * s_stage_bytes(R);
    static bool attr_set[64] = {};
    int dev = 0;
    if (cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaFuncSetAttribute(warp_bwd_cell_staged_kernel<R, GF, MB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_set[dev] = true;
    }
    warp_bwd_cell_staged_kernel<R, GF, MB><<<grid, S_THREADS, smem, st>>>(tm_gs, tm_gl, tm_in, gout, input, gin, gflow, tab, wts,
                                                                        counts, bent, p, nbx, nby, csplit, c_per);
    return cudaGetLastError();
}

// -1: shape not eligible (the caller falls back); otherwise a cudaError_t.
int dsvc_warp_bwd_cell_launch(const float* gout, const float* input, const float* flow, float* gin, float* gflow,
                              const float* lin_x, const float* lin_y, const WarpParams& p, void* workspace,
                              size_t workspace_bytes, cudaStream_t st) {
    using namespace bcell;
    if (!gin || !workspace || !aligned16(workspace)) return -1;
    if ((long long)p.H * p.W >= (1ll << 30) || (long long)p.B * p.H * p.W >= (1ll << 31)) return -1;
    if (p.H > 65535 || p.W > 65535 || p.B > 65535) return -1;
    const int R = cell_rows(p);
    const Layout L = make_layout(p.B, p.H, p.W, R);
    if (workspace_bytes < L.total) return -1;
    char* ws = static_cast<char*>(workspace);
    int* counts = reinterpret_cast<int*>(ws + L.counts_off);
    int* tab = reinterpret_cast<int*>(ws + L.tab_off);
    float2* wts = reinterpret_cast<float2*>(ws + L.wts_off);
    int4* bent = reinterpret_cast<int4*>(ws + L.bent_off);
    int2* ovf = reinterpret_cast<int2*>(ws + L.ovf_off);
    // one fill: the counters (they count from -1) and the tables behind them (-1 = empty layer)
    cudaError_t e = cudaMemsetAsync(counts, 0xFF, L.counts_bytes + L.tab_bytes, st);
    if (e != cudaSuccess) return (int)e;

#ifdef DSVC_TUNE
    static const int minb_k = cell_knob("DSVC_CELL_MINB", 0), ring_k = cell_knob("DSVC_CELL_RING", -1);
    const int minb = minb_k ? minb_k : (R == 2 ? 3 : 2), use_ring = ring_k >= 0 ? ring_k : (R == 3 ? 1 : 0);
#else
    const int minb = 3;
#endif
    // channel ranges only when the regions alone leave SMs idle (grad_flow then accumulates)
    const long long ctas = (long long)L.nrx * L.nry * p.B;
    const long long resident = (long long)minb * DSVC_NUM_SMS;
    int csplit = 1;
    if (ctas < 4 * resident) {
        // pick the split (1, 2, 4) whose last wave is fullest
        double best = -1.0;
        for (int s = 1; s <= 4 && s <= p.C; s *= 2) {
            const double waves = (double)(ctas * s) / (double)resident;
            const double eff = waves / (double)(long long)(waves + 0.999999);
            if (eff > best + 0.02) { best = eff; csplit = s; }
        }
    }
    const int c_per = (p.C + csplit - 1) / csplit;
    if (gflow && csplit > 1) {
        e = cudaMemsetAsync(gflow, 0, (size_t)p.B * 2 * p.H * p.W * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
    }
    // (No programmatic dependent launch here: the kernels read what their predecessor wrote through
    // the read-only path, which must not overlap the writer's lifetime -- measured gain 2 %, not taken.)
    cell_build_kernel<<<dim3((p.W + 255) / 256, p.H, p.B), 256, 0, st>>>(flow, lin_x, lin_y, p, tab, wts, counts, bent, ovf,
                                                                         L.nrx, L.nry, CTA_BY * R);
    const dim3 grid(L.nrx, L.nry, p.B * csplit);
#define DSVC_CELL_GO(RR, MB, RG)                                                                                                  \
    e = gflow ? cell_launch_one<RR, true, MB, RG>(grid, st, gout, input, gin, gflow, tab, wts, counts, bent, p, L.nbx, L.nby, csplit, c_per) \
              : cell_launch_one<RR, false, MB, RG>(grid, st, gout, input, gin, gflow, tab, wts, counts, bent, p, L.nbx, L.nby, csplit, c_per)
#define DSVC_CELL_GO_STAGED(RR, MB)                                                                                               \
    e = gflow ? cell_launch_staged<RR, true, MB>(grid, st, gout, input, gin, gflow, tab, wts, counts, bent, p, L.nbx, L.nby, csplit, c_per) \
              : cell_launch_staged<RR, false, MB>(grid, st, gout, input, gin, gflow, tab, wts, counts, bent, p, L.nbx, L.nby, csplit, c_per)
    // TMA needs 16-byte rows
    const bool can_stage = p.W % 4 == 0 && aligned16(gout) && aligned16(input) && (long long)p.B * p.C < (1ll << 30);
#ifdef DSVC_TUNE
    static const int staged_k = cell_knob("DSVC_CELL_STAGED", 1);
#else
    const int staged_k = 1;
#endif
    e = cudaErrorNotSupported;
    if (can_stage && staged_k) {
#ifdef DSVC_TUNE
        if (R == 3 && minb == 3) DSVC_CELL_GO_STAGED(3, 3);
        else if (R == 3) DSVC_CELL_GO_STAGED(3, 2);
        else if (R == 4) DSVC_CELL_GO_STAGED(4, 2);
        else if (minb == 2) DSVC_CELL_GO_STAGED(2, 2);
        else
#endif
        DSVC_CELL_GO_STAGED(2, DSVC_CELL_STAGED_MINB);
    }
    if (e == cudaErrorNotSupported) {  // unaligned rows or no tensor-map encoder: the load variant
#ifdef DSVC_TUNE
        if (R == 4 && minb == 2 && use_ring) DSVC_CELL_GO(4, 2, true);
        else if (R == 4 && minb == 2) DSVC_CELL_GO(4, 2, false);
        else if (R == 4 && use_ring) DSVC_CELL_GO(4, 3, true);
        else if (R == 4) DSVC_CELL_GO(4, 3, false);
        else if (R == 3 && minb == 4) DSVC_CELL_GO(3, 4, false);
        else if (R == 3 && use_ring) DSVC_CELL_GO(3, 3, true);
        else if (R == 3) DSVC_CELL_GO(3, 3, false);
        else if (use_ring) DSVC_CELL_GO(2, 4, true);
        else
#endif
        DSVC_CELL_GO(2, 4, false);
    }
#undef DSVC_CELL_GO_STAGED
#undef DSVC_CELL_GO
    if (e != cudaSuccess) return (int)e;
    if (gflow)
        cell_fixup_kernel<true><<<2 * DSVC_NUM_SMS, 256, 0, st>>>(gout, input, flow, gin, gflow, lin_x, lin_y, p, counts, ovf,
                                                                  CTA_BY * R);
    else
        cell_fixup_kernel<false><<<2 * DSVC_NUM_SMS, 256, 0, st>>>(gout, input, flow, gin, gflow, lin_x, lin_y, p, counts, ovf,
                                                                   CTA_BY * R);
    return (int)cudaGetLastError();
}
