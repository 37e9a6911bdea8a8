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
// Launches of one call (dsvc_warp_bwd_cell_launch):
//   1. cell_build_kernel, one thread per output pixel: coordinates as in the forward (bit-exact
//      op order, warp_bwd_common.cuh), then the pixel claims layer 0 or 1 of its cell's table
//      entry by atomicCAS (pixel index + clamp flags, weights beside it).  A third, fourth ...
//      pixel of a cell is an EXTRA: it is appended to the bucket (<= 256 entries) of every
//      62 x 4R region of grad_input (one CTA of the next launch) that one of its taps reaches;
//      with a full bucket it goes to the launch-wide overflow list with the mask of the regions
//      that refused it.
//   2. warp_bwd_cell_kernel: one warp per 31 x R block of grad_input (+1 halo column / row of
//      cells), lane = cell column; two table layers in registers; per channel 2(R+1) loads of
//      grad_out, R+1 row loads of the input, 3R+2 shuffles, R row stores.  After the CTA's warps
//      have stored their blocks, one thread per bucket entry walks the channels and adds the
//      extra's taps inside the region with RED.ADD (and computes its grad_flow).
//   3. cell_fixup_kernel: the overflow list (empty for SpyNet-like flows), one warp per entry,
//      per-pixel scatter of exactly the taps nobody else took.
// Any flow is handled; the speed degrades towards the per-pixel kernel as cells fill up.
#include "warp_bwd_common.cuh"
#include <cstdlib>

namespace dsvc {
namespace bcell {

constexpr int DC = 31;  // columns of grad_input per warp (lane 0 is the halo column of cells)
constexpr int WARPS = 8, THREADS = WARPS * 32;
constexpr int CTA_BX = 2, CTA_BY = 4;  // a CTA's warps cover a region of 62 x 4R elements
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
    size_t counts_off, counts_bytes;  // int32: [0] overflow entries, [4 + region] bucket fill (zeroed per call)
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
        const int k = atomicAdd(counts + 4 + reg, 1);
        if (k < EC) bent[(size_t)reg * EC + k] = ent;
        else fail |= bit;
    };
    put(rx, ry, 1);
    if (ex) put(rx + 1, ry, 2);
    if (ey) put(rx, ry + 1, 4);
    if (ex && ey) put(rx + 1, ry + 1, 8);
    if (fail) {
        const int k = atomicAdd(counts, 1);
        ovf[k] = make_int2((int)((size_t)b * plane + pix), fail);
    }
}

// ---------------------------------------------------------------------------------- 2. the gradients
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
    const float mx = (float)(W - 1) * 0.5f, my = (float)(H - 1) * 0.5f;
    // grad_flow = clip multiplier * channel sum / scale  (store_gflow, warp_bwd_common.cuh)
    auto put_gflow = [&](unsigned pix, unsigned cl, float sx_, float sy_) {
        float* gf = gflow + (size_t)b * 2 * plane;
        const float ggx = __fmul_rn((cl & 1u) ? 0.0f : mx, sx_), ggy = __fmul_rn((cl & 2u) ? 0.0f : my, sy_);
        const float vx = p.flow_mode ? __fdiv_rn(ggx, p.sx) : __fmul_rn(ggx, p.inv_sx);
        const float vy = p.flow_mode ? __fdiv_rn(ggy, p.sy) : __fmul_rn(ggy, p.inv_sy);
        if (csplit > 1) {
            atomicAdd(gf + pix, vx);
            atomicAdd(gf + plane + pix, vy);
        } else {
            gf[pix] = vx;
            gf[plane + pix] = vy;
        }
    };

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

            // per cell: S0 = sum g, Sx = sum g wx1, Sy = sum g wy1, Sxy = sum g wx1 wy1 over its pixels;
            // south-east share = Sxy, south-west = Sy - Sxy, north-east = Sx - Sxy, north-west = the rest
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
                        put_gflow(off[lr][e], word >> 30, gx[lr - 1][e], gy[lr - 1][e]);
                    }
            }
        }
    }

    // ---- the region's extras: one thread per bucket entry, behind every warp's stores
    const int region = (b * (int)gridDim.y + (int)blockIdx.y) * (int)gridDim.x + (int)blockIdx.x;
    const int n_ex = min(__ldg(counts + 4 + region), EC);
    if (n_ex == 0) return;  // CTA-uniform
    __syncthreads();
    if ((int)threadIdx.x >= n_ex || c0 >= c1) return;
    const int4 e = __ldg(bent + (size_t)region * EC + threadIdx.x);
    const unsigned pix = (unsigned)e.x & PIX_MASK;
    const float ewx = __int_as_float(e.y), ewy = __int_as_float(e.z);
    const int cx = e.w & 0xffff, cy = (int)((unsigned)e.w >> 16);
    constexpr int RY = CTA_BY * R;
    const int rx = (int)blockIdx.x, ry = (int)blockIdx.y;
    const bool xe = cx + 1 < W, ys = cy + 1 < H;
    const bool cx0 = cx / RX == rx, cx1 = xe && (cx + 1) / RX == rx, cy0 = cy / RY == ry, cy1 = ys && (cy + 1) / RY == ry;
    const bool t0 = cx0 && cy0, t1 = cx1 && cy0, t2 = cx0 && cy1, t3 = cx1 && cy1;
    const bool own = NEED_GFLOW && t0;  // the region of the cell's own element computes the pixel's grad_flow
    const unsigned ebase = (unsigned)(cy * W + cx);
    const float w_se = ewx * ewy, w_sw = ewy - w_se, w_ne = ewx - w_se, w_nw = (1.0f - ewx) - w_sw;
    float egx = 0.0f, egy = 0.0f;
    const float* gp = gout + ((size_t)b * p.C + c0) * plane + pix;
    const float* ip = in + ((size_t)b * p.C + c0) * plane + ebase;
    float* op = gin + ((size_t)b * p.C + c0) * plane + ebase;
    constexpr int U = 4;
    for (int c = c0; c < c1; c += U) {
        float ge[U], v00[U], v01[U], v10[U], v11[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const bool cok = c + u < c1;
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
            const bool cok = c + u < c1;
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
    if (own) put_gflow(pix, (unsigned)e.x >> 30, egx, egy);
}

// ---------------------------------------------------------------------------------- 3. overflow
template <bool NEED_GFLOW>
__global__ void __launch_bounds__(256)
cell_fixup_kernel(const float* __restrict__ gout, const float* __restrict__ in, const float* __restrict__ flow,
                  float* __restrict__ gin, float* __restrict__ gflow, const float* __restrict__ lin_x,
                  const float* __restrict__ lin_y, WarpParams p, const int* __restrict__ counts,
                  const int2* __restrict__ ovf, int RY) {
    const int n = __ldg(counts);
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
static int cell_rows() {
#ifdef DSVC_TUNE
    static const int r = cell_knob("DSVC_CELL_R", 3);
    return (r == 2 || r == 4) ? r : 3;
#else
    return 3;
#endif
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
    if (smem > 48 * 1024 && cudaGetDevice(&dev) == cudaSuccess && dev >= 0 && dev < 64 && !attr_set[dev]) {
        cudaFuncSetAttribute(warp_bwd_cell_kernel<R, GF, MB, RING>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
        attr_set[dev] = true;
    }
    warp_bwd_cell_kernel<R, GF, MB, RING><<<grid, THREADS, smem, st>>>(gout, input, gin, gflow, tab, wts, counts, bent, p, nbx, nby,
                                                                     csplit, c_per);
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
    const int R = cell_rows();
    const Layout L = make_layout(p.B, p.H, p.W, R);
    if (workspace_bytes < L.total) return -1;
    char* ws = static_cast<char*>(workspace);
    int* counts = reinterpret_cast<int*>(ws + L.counts_off);
    int* tab = reinterpret_cast<int*>(ws + L.tab_off);
    float2* wts = reinterpret_cast<float2*>(ws + L.wts_off);
    int4* bent = reinterpret_cast<int4*>(ws + L.bent_off);
    int2* ovf = reinterpret_cast<int2*>(ws + L.ovf_off);
    cudaError_t e = cudaMemsetAsync(counts, 0, L.counts_bytes, st);
    if (e != cudaSuccess) return (int)e;
    e = cudaMemsetAsync(tab, 0xFF, L.tab_bytes, st);
    if (e != cudaSuccess) return (int)e;

#ifdef DSVC_TUNE
    static const int minb = cell_knob("DSVC_CELL_MINB", 3), use_ring = cell_knob("DSVC_CELL_RING", 0);
#else
    const int minb = 3, use_ring = 0;
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
    cell_build_kernel<<<dim3((p.W + 255) / 256, p.H, p.B), 256, 0, st>>>(flow, lin_x, lin_y, p, tab, wts, counts, bent, ovf,
                                                                         L.nrx, L.nry, CTA_BY * R);
    const dim3 grid(L.nrx, L.nry, p.B * csplit);
#define DSVC_CELL_GO(RR, MB, RG)                                                                                                  \
    e = gflow ? cell_launch_one<RR, true, MB, RG>(grid, st, gout, input, gin, gflow, tab, wts, counts, bent, p, L.nbx, L.nby, csplit, c_per) \
              : cell_launch_one<RR, false, MB, RG>(grid, st, gout, input, gin, gflow, tab, wts, counts, bent, p, L.nbx, L.nby, csplit, c_per)
#ifdef DSVC_TUNE
    if (R == 4 && minb == 2 && use_ring) DSVC_CELL_GO(4, 2, true);
    else if (R == 4 && minb == 2) DSVC_CELL_GO(4, 2, false);
    else if (R == 4 && use_ring) DSVC_CELL_GO(4, 3, true);
    else if (R == 4) DSVC_CELL_GO(4, 3, false);
    else if (R == 3 && minb == 4) DSVC_CELL_GO(3, 4, false);
    else if (R == 3 && use_ring) DSVC_CELL_GO(3, 3, true);
    else if (R == 3) DSVC_CELL_GO(3, 3, false);
    else if (use_ring) DSVC_CELL_GO(2, 4, true);
    else DSVC_CELL_GO(2, 4, false);
#else
    DSVC_CELL_GO(3, 3, false);
#endif
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
