// Backward bilinear warp (motion compensation), direct-gather kernels.
//
// Replaces /root/reference/modules.py:25-62 (torch_warp) and its autograd.
// These kernels read the four taps straight through the read-only path; the
// shared-memory/TMA-staged forward lives in warp_persist.cu and is preferred when the
// shape allows.  HBM-bound op: one launch reads input + flow once and writes out
// once (4*B*H*W*(2C+2) algorithmic bytes); coordinates and weights are computed
// once per pixel and reused for every channel.
#include <algorithm>
#include <type_traits>
#include <cstdlib>

#include "warp_bwd_common.cuh"
#include "warp_common.cuh"
#include "../../include/deepsvc_b200.h"

namespace dsvc {

// ------------------------------------------------------------------ NCHW forward
// CTA = 32 x 8 output pixels (one warp per 128-byte row segment, so that the
// y/y+1 taps of neighbouring warps hit the same L1 lines), channels [c0, c0+cpc).
template <int UNROLL>
__global__ void __launch_bounds__(256)
warp_fwd_nchw_gather(const float* __restrict__ in, const float* __restrict__ flow,
                     float* __restrict__ out, const float* __restrict__ lin_x,
                     const float* __restrict__ lin_y, WarpParams p, int cpc, int nchunk) {
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    const int b = blockIdx.z / nchunk;
    const int c0 = (blockIdx.z - b * nchunk) * cpc;
    if (x >= p.W || y >= p.H) return;
    gather_pixel<UNROLL>(in, flow, out, lin_x, lin_y, p, b, x, y, c0, min(p.C, c0 + cpc));
}

// ------------------------------------------------------------------ NCHW forward, few channels
// Frames and SpyNet pyramid levels (C = 3; modules.py:167, video_model.py:37).  With so few
// channels the per-pixel coordinate and address arithmetic is the cost (the one-pixel-per-
// thread gather issues ~250 instructions per pixel and is issue-bound at 2.7 TB/s), so this
// kernel keeps it lean: 32-bit element offsets, taps outside the image read a clamped
// in-image address with an exactly-zero weight instead of being branched around, a
// persistent grid works through 32 x (8*PX) pixel tiles, every thread keeps the 4*C*PX taps of
// its PX pixels in flight at once and loads the NEXT tile's flow before it gathers the
// current one.  Results are identical to gather_pixel's on finite inputs.
// Tiles are CLAIMED from a device counter when the caller provides scheduler state (`sched`, the
// same zero-before / zero-after words as the persistent staged kernel): next to the feature warp
// only a fraction of this grid's CTAs is resident at once, and a static tile assignment left the
// late CTAs' tiles to run after it (round 1: the 3-ch warps added their whole time to the frame).
template <int C, int PX, int FM>
__global__ void __launch_bounds__(256)
warp_fwd_nchw_fewch(const float* __restrict__ in, const float* __restrict__ flow,
                    float* __restrict__ out, const float* __restrict__ lin_x,
                    const float* __restrict__ lin_y, WarpParams p, unsigned tiles_x,
                    unsigned tiles_y, unsigned total_tiles, WarpSched* __restrict__ sched) {
    const unsigned lane = threadIdx.x, wy = threadIdx.y;
    const unsigned W = p.W, H = p.H, plane = H * W;  // C*H*W < 2^31 (checked by the launcher)
    float fx[PX], fy[PX];
    __shared__ uint4 s_claim[2];   // claimed tile: (index, tx, ty, b), decoded once by the claiming thread
    pdl_prologue();
    struct Tile { unsigned tx, ty, b; };
    auto decode = [&](unsigned t, Tile& q) {
        const unsigned r = t / tiles_x;
        q.tx = t - r * tiles_x;
        q.b = r / tiles_y;
        q.ty = r - q.b * tiles_y;
    };
    // static order (no scheduler state): tile coordinates advance by gridDim.x without divisions
    const unsigned step_r = gridDim.x / tiles_x, step_x = gridDim.x - step_r * tiles_x;
    const unsigned step_b = step_r / tiles_y, step_y = step_r - step_b * tiles_y;
    unsigned n_claims = 0;
    // the tile after (t, q): the next one off the counter, or the grid stride
    auto next_tile = [&](unsigned t, Tile& q) -> unsigned {
        if (!sched) {
            q.tx += step_x;
            const unsigned cx = q.tx >= tiles_x ? 1u : 0u;
            q.tx -= cx ? tiles_x : 0u;
            q.ty += step_y + cx;
            const unsigned cy = q.ty >= tiles_y ? 1u : 0u;
            q.ty -= cy ? tiles_y : 0u;
            q.b += step_b + cy;
            return t + gridDim.x;
        }
        const unsigned slot = n_claims & 1u;
        if (lane == 0 && wy == 0) {
            const unsigned c = (unsigned)atomicAdd(&sched->next, 1);
            Tile d;
            decode(c, d);
            s_claim[slot] = make_uint4(c, d.tx, d.ty, d.b);
        }
        __syncthreads();  // (two slots: a warp still reading the previous claim is never overwritten)
        ++n_claims;
        const uint4 c = s_claim[slot];
        q.tx = c.y; q.ty = c.z; q.b = c.w;
        return c.x;
    };
    auto load_flow = [&](const Tile& q) {
        const unsigned x = q.tx * 32 + lane, y0 = q.ty * (8 * PX) + wy;
        const float* fl = flow + (size_t)q.b * 2 * plane;
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            const unsigned y = y0 + 8 * j;
            const unsigned pix = (x < W && y < H) ? y * W + x : 0u;
            fx[j] = __ldg(fl + pix);
            fy[j] = __ldg(fl + (pix + plane));
        }
    };
    Tile next;
    unsigned t;
    if (sched) {
        t = next_tile(0u, next);
    } else {
        t = blockIdx.x;
        decode(t, next);
    }
    if (t < total_tiles) load_flow(next);
    while (t < total_tiles) {
        const Tile cur = next;
        const unsigned b = cur.b, x = cur.tx * 32 + lane, y0 = cur.ty * (8 * PX) + wy;
        float cfx[PX], cfy[PX];
#pragma unroll
        for (int j = 0; j < PX; ++j) { cfx[j] = fx[j]; cfy[j] = fy[j]; }
        t = next_tile(t, next);
        if (t < total_tiles) load_flow(next);  // in flight during this tile
        const float* inb = in + (size_t)b * C * plane;
        float* outb = out + (size_t)b * C * plane;
        const float lx = __ldg(lin_x + min(x, W - 1));
        float wnw[PX], wne[PX], wsw[PX], wse[PX];
        const float* qn[PX];
        bool ok[PX];
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            const unsigned y = y0 + 8 * j;
            ok[j] = x < W && y < H;
            const float ix = source_coord(lx, cfx[j], p.sx, p.inv_sx, FM, (int)W);
            const float iy = source_coord(__ldg(lin_y + min(y, H - 1)), cfy[j], p.sy, p.inv_sy, FM, (int)H);
            const Taps tp = make_taps(ix, iy, (int)W, (int)H);
            // The 2 x 2 footprint always starts at (min(x0, W-2), min(y0, H-2)), so that the four
            // taps sit at fixed offsets (+1, +W) from one address.  A coordinate on the far border
            // (x0 = W-1, where the reference skips the outside taps and their weights are exactly
            // 0) moves its weights one column / row over instead: same products, same order.
            const bool sx_ = !tp.x1ok, sy_ = !tp.y1ok;
            const unsigned xb = sx_ ? W - 2 : (unsigned)tp.x0, yb = sy_ ? H - 2 : (unsigned)tp.y0;
            const float n0 = sx_ ? 0.0f : tp.nw, n1 = sx_ ? tp.nw : tp.ne;  // north row, west / east
            const float s0 = sx_ ? 0.0f : tp.sw, s1 = sx_ ? tp.sw : tp.se;  // south row
            wnw[j] = sy_ ? 0.0f : n0;
            wne[j] = sy_ ? 0.0f : n1;
            wsw[j] = sy_ ? n0 : s0;
            wse[j] = sy_ ? n1 : s1;
            qn[j] = inb + (ok[j] ? yb * W + xb : 0u);
        }
        float v[PX][C][4];
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            const float* q = qn[j];
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const float* qs = q + W;
                v[j][c][0] = __ldg(q);
                v[j][c][1] = __ldg(q + 1);
                v[j][c][2] = __ldg(qs);
                v[j][c][3] = __ldg(qs + 1);
                q += plane;
            }
        }
#pragma unroll
        for (int j = 0; j < PX; ++j) {
            if (!ok[j]) continue;
            float* op = outb + ((y0 + 8 * j) * W + x);
#pragma unroll
            for (int c = 0; c < C; ++c) {
                float acc = __fmul_rn(v[j][c][0], wnw[j]);
                acc = fmaf(v[j][c][1], wne[j], acc);
                acc = fmaf(v[j][c][2], wsw[j], acc);
                acc = fmaf(v[j][c][3], wse[j], acc);
                st_stream1(op, acc);
                op += plane;
            }
        }
    }
    if (sched && lane == 0 && wy == 0) {
        // the last CTA to run dry leaves the scheduler state zeroed for the next launch
        if (atomicAdd(&sched->exited, 1) == (int)gridDim.x - 1) {
            sched->next = 0;
            __threadfence();
            sched->exited = 0;
        }
    }
}

// ------------------------------------------------------------------ NHWC forward
// channels_last: every tap is C contiguous floats -> pure 128-bit loads.  G lanes
// cooperate on one pixel (G*16 bytes per tap per instruction), C % 4 == 0.
template <int G>
__global__ void __launch_bounds__(256)
warp_fwd_nhwc(const float* __restrict__ in, const float* __restrict__ flow,
              float* __restrict__ out, const float* __restrict__ lin_x,
              const float* __restrict__ lin_y, WarpParams p) {
    const size_t plane = (size_t)p.H * p.W;
    const size_t gpix = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) / G;  // b*H*W + y*W + x
    const int sub = threadIdx.x % G;
    if (gpix >= (size_t)p.B * plane) return;
    const int b = (int)(gpix / plane);
    const size_t pix = gpix - (size_t)b * plane;
    const int y = (int)(pix / p.W), x = (int)(pix - (size_t)y * p.W);
    const float* fl = flow + (size_t)b * 2 * plane + pix;
    const float fx = __ldg(fl), fy = __ldg(fl + plane);
    const float ix = source_coord(__ldg(lin_x + x), fx, p.sx, p.inv_sx, p.flow_mode, p.W);
    const float iy = source_coord(__ldg(lin_y + y), fy, p.sy, p.inv_sy, p.flow_mode, p.H);
    const Taps t = make_taps(ix, iy, p.W, p.H);
    const int C4 = p.C >> 2;
    const float4* base = reinterpret_cast<const float4*>(in) + (size_t)b * plane * C4;
    const float4* q_nw = base + ((size_t)t.y0 * p.W + t.x0) * C4;
    const size_t dx = t.x1ok ? (size_t)C4 : 0, dy = t.y1ok ? (size_t)p.W * C4 : 0;
    float4* o = reinterpret_cast<float4*>(out) + gpix * C4;
    const bool xe = t.x1ok, ys = t.y1ok, xy = t.x1ok && t.y1ok;
    for (int c4 = sub; c4 < C4; c4 += G) {
        const float4 a = __ldg(q_nw + c4);
        const float4 bq = __ldg(q_nw + dx + c4);
        const float4 cq = __ldg(q_nw + dy + c4);
        const float4 d = __ldg(q_nw + dy + dx + c4);
        float4 r;
#define DSVC_TAP4(f)                                   \
    r.f = __fmul_rn(a.f, t.nw);                        \
    r.f = xe ? fmaf(bq.f, t.ne, r.f) : r.f;            \
    r.f = ys ? fmaf(cq.f, t.sw, r.f) : r.f;            \
    r.f = xy ? fmaf(d.f, t.se, r.f) : r.f;
        DSVC_TAP4(x) DSVC_TAP4(y) DSVC_TAP4(z) DSVC_TAP4(w)
#undef DSVC_TAP4
        st_stream4(reinterpret_cast<float*>(o + c4), r);
    }
}

// ------------------------------------------------------------------ backward (NCHW)
// One thread per output pixel, all channels (warp_bwd_common.cuh): the general-shape kernel.
// The bandwidth-critical calls (C >= 8, 16-byte aligned rows) take the staged kernel of
// warp_bwd_staged.cu, which pre-combines the scatter per tile in shared memory.
template <bool NEED_GIN, bool NEED_GFLOW>
__global__ void __launch_bounds__(256)
warp_bwd_nchw(const float* __restrict__ gout, const float* __restrict__ in,
              const float* __restrict__ flow, float* __restrict__ gin,
              float* __restrict__ gflow, const float* __restrict__ lin_x,
              const float* __restrict__ lin_y, WarpParams p, const int* __restrict__ mode) {
    if (mode && *mode == DSVC_BWD_MODE_STAGED) return;  // the staged kernel ahead of this launch did the job
    const int x = blockIdx.x * 32 + threadIdx.x;
    const int y = blockIdx.y * 8 + threadIdx.y;
    if (x >= p.W || y >= p.H) return;
    bwd_pixel_direct<NEED_GIN, NEED_GFLOW>(gout, in, flow, gin, gflow, lin_x, lin_y, p, blockIdx.z, x, y,
                                           0, p.C, false);
}

}  // namespace dsvc

using namespace dsvc;

int dsvc_warp_fwd_persist_launch(const float* input, const float* flow, float* out,
                                 const float* lin_x, const float* lin_y, const WarpParams& p,
                                 bool force, void* workspace, size_t workspace_bytes,
                                 cudaStream_t st, const float* input2 = nullptr, float* out2 = nullptr,
                                 int C2 = 0);  // warp_persist.cu

int dsvc_warp_bwd_staged_launch(const float* gout, const float* input, const float* flow, float* gin,
                                float* gflow, const float* lin_x, const float* lin_y, const WarpParams& p,
                                bool force, cudaStream_t st, const int* mode = nullptr);  // warp_bwd_staged.cu

static int warp_args_ok(const void* a, const void* b, const void* c, int B, int C, int H, int W,
                        const void* lx, const void* ly) {
    return a && b && c && lx && ly && B > 0 && C > 0 && H > 0 && W > 0 &&
           (int64_t)H * W < (int64_t)1 << 30 && B <= 65535;
}

extern "C" int dsvc_warp_fwd_f32(const float* input, const float* flow, float* out, int B, int C,
                                 int H, int W, const float* lin_x, const float* lin_y, float sx,
                                 float sy, float inv_sx, float inv_sy, int flow_mode, int layout,
                                 int algo, void* workspace, size_t workspace_bytes,
                                 void* stream) {
    DSVC_CHECK_ARG(warp_args_ok(input, flow, out, B, C, H, W, lin_x, lin_y));
    DSVC_CHECK_ARG(flow_mode == 0 || flow_mode == 1);
    cudaStream_t st = (cudaStream_t)stream;
    WarpParams p{B, C, H, W, sx, sy, inv_sx, inv_sy, flow_mode};
    if (layout == DSVC_LAYOUT_NHWC) {
        DSVC_CHECK_ARG(C % 4 == 0 && aligned16(input) && aligned16(out));
        DSVC_CHECK_ARG(algo != DSVC_WARP_TMA);
        const int C4 = C / 4;
        const int G = C4 >= 8 ? 8 : (C4 >= 4 ? 4 : (C4 >= 2 ? 2 : 1));
        const size_t nthreads = (size_t)B * H * W * G;
        const unsigned grid = (unsigned)((nthreads + 255) / 256);
        switch (G) {
            case 8: warp_fwd_nhwc<8><<<grid, 256, 0, st>>>(input, flow, out, lin_x, lin_y, p); break;
            case 4: warp_fwd_nhwc<4><<<grid, 256, 0, st>>>(input, flow, out, lin_x, lin_y, p); break;
            case 2: warp_fwd_nhwc<2><<<grid, 256, 0, st>>>(input, flow, out, lin_x, lin_y, p); break;
            default: warp_fwd_nhwc<1><<<grid, 256, 0, st>>>(input, flow, out, lin_x, lin_y, p); break;
        }
        DSVC_RETURN_LAST();
    }
    DSVC_CHECK_ARG(layout == DSVC_LAYOUT_NCHW);
    if (algo == DSVC_WARP_AUTO || algo == DSVC_WARP_TMA) {
        const int r = dsvc_warp_fwd_persist_launch(input, flow, out, lin_x, lin_y, p, algo == DSVC_WARP_TMA,
                                                   workspace, workspace_bytes, st);
        if (r != -1) return r;  // -1: shape not supported by the staged kernel -> gather
        if (algo == DSVC_WARP_TMA) return (int)cudaErrorInvalidValue;
    }
    if (C <= 4 && H >= 2 && W >= 2 && (long long)C * H * W < (1ll << 31)) {
        // frames / pyramid levels: persistent few-channel kernel, one CTA per resident slot
        {
            constexpr int PX = 2;  // pixels per thread (1, 3, 4 measured slower: 1080p C=3 19.4 us at 2;
                                   // 64 x 8 and 128 x 8 row-major tiles measured the same as 32 x 16)
            const unsigned tiles_x = (W + 31) / 32, tiles_y = (H + 8 * PX - 1) / (8 * PX);
            const long long total = (long long)tiles_x * tiles_y * B;
            if (total >= (1ll << 31)) return (int)cudaErrorInvalidValue;
            auto launch = [&](auto kernel) -> int {
                static int slots = 0;  // per instantiation (all devices of a box are the same part)
                if (slots == 0) {
                    int dev = 0, sms = 0, per_sm = 0;
                    if (cudaGetDevice(&dev) != cudaSuccess ||
                        cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0)
                        sms = DSVC_NUM_SMS;
                    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, 256, 0) != cudaSuccess || per_sm <= 0)
                        per_sm = 2;
                    slots = sms * per_sm;
                }
                // (no carve-out preference here: this kernel lives on L1 hits -- 26.8 -> 32.8 us at 1080p
                // with the maximum shared-memory carve-out; asking for it does not change the frame's DAG
                // time either: 256.8 vs 257.5 us, round 2)
                const int grid = (int)std::min<long long>(total, slots);
                WarpSched* sched = (workspace && workspace_bytes >= sizeof(WarpSched) && aligned16(workspace))
                                       ? static_cast<WarpSched*>(workspace) : nullptr;
                launch_pdl(kernel, dim3(grid), dim3(32, 8), 0, st, input, flow, out, lin_x, lin_y, p, tiles_x, tiles_y,
                           (unsigned)total, sched);
                return (int)cudaGetLastError();
            };
#define DSVC_FEWCH(CN) return flow_mode ? launch(warp_fwd_nchw_fewch<CN, PX, 1>) : launch(warp_fwd_nchw_fewch<CN, PX, 0>)
            switch (C) {
                case 1: DSVC_FEWCH(1);
                case 2: DSVC_FEWCH(2);
                case 3: DSVC_FEWCH(3);
                default: DSVC_FEWCH(4);
            }
#undef DSVC_FEWCH
        }
    }
    // channel chunk per CTA: enough CTAs to fill the machine, few flow re-reads
    int cpc = C;
    if (C > 16) cpc = 16;
    const int nchunk = (C + cpc - 1) / cpc;
    DSVC_CHECK_ARG((int64_t)B * nchunk <= 65535);
    dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8, B * nchunk);
    // all taps of up to 4 channels in flight per thread (C = 3 frames: 12 loads at once)
    if (cpc >= 4)
        warp_fwd_nchw_gather<4><<<grid, block, 0, st>>>(input, flow, out, lin_x, lin_y, p, cpc, nchunk);
    else if (cpc == 3)
        warp_fwd_nchw_gather<3><<<grid, block, 0, st>>>(input, flow, out, lin_x, lin_y, p, cpc, nchunk);
    else if (cpc == 2)
        warp_fwd_nchw_gather<2><<<grid, block, 0, st>>>(input, flow, out, lin_x, lin_y, p, cpc, nchunk);
    else
        warp_fwd_nchw_gather<1><<<grid, block, 0, st>>>(input, flow, out, lin_x, lin_y, p, cpc, nchunk);
    DSVC_RETURN_LAST();
}

extern "C" int dsvc_warp_fwd2_f32(const float* input_a, const float* input_b, const float* flow,
                                  float* out_a, float* out_b, int B, int Ca, int Cb, int H, int W,
                                  const float* lin_x, const float* lin_y, float sx, float sy,
                                  float inv_sx, float inv_sy, int flow_mode, void* workspace,
                                  size_t workspace_bytes, void* stream) {
    DSVC_CHECK_ARG(warp_args_ok(input_a, flow, out_a, B, Ca, H, W, lin_x, lin_y));
    DSVC_CHECK_ARG(input_b && out_b && Cb > 0);
    DSVC_CHECK_ARG(flow_mode == 0 || flow_mode == 1);
    WarpParams p{B, Ca + Cb, H, W, sx, sy, inv_sx, inv_sy, flow_mode};
    if (Ca >= 8 && W >= 64 && H >= 32) {
        const int r = dsvc_warp_fwd_persist_launch(input_a, flow, out_a, lin_x, lin_y, p, false, workspace,
                                                   workspace_bytes, (cudaStream_t)stream, input_b, out_b, Cb);
        if (r != -1) return r;
    }
    // shape not eligible for the shared launch: two ordinary launches (same results)
    const int r = dsvc_warp_fwd_f32(input_a, flow, out_a, B, Ca, H, W, lin_x, lin_y, sx, sy, inv_sx, inv_sy,
                                    flow_mode, DSVC_LAYOUT_NCHW, DSVC_WARP_AUTO, workspace, workspace_bytes, stream);
    if (r) return r;
    return dsvc_warp_fwd_f32(input_b, flow, out_b, B, Cb, H, W, lin_x, lin_y, sx, sy, inv_sx, inv_sy, flow_mode,
                             DSVC_LAYOUT_NCHW, DSVC_WARP_AUTO, workspace, workspace_bytes, stream);
}

extern "C" size_t dsvc_warp_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return sizeof(WarpSched) + 48;  // scheduler state of the persistent staged kernel
}

static int g_bwd_algo = 0;  // DSVC_WARP_BWD_* (dsvc_set_warp_bwd_algo: tests / profiling)
extern "C" int dsvc_set_warp_bwd_algo(int algo) {
    DSVC_CHECK_ARG(algo >= 0 && algo <= 4);
    g_bwd_algo = algo;
    return 0;
}

extern "C" int dsvc_warp_bwd_f32(const float* grad_out, const float* input, const float* flow,
                                 float* grad_input, float* grad_flow, int B, int C, int H, int W,
                                 const float* lin_x, const float* lin_y, float sx, float sy,
                                 float inv_sx, float inv_sy, int flow_mode, int layout,
                                 void* stream) {
    DSVC_CHECK_ARG(warp_args_ok(grad_out, input, flow, B, C, H, W, lin_x, lin_y));
    DSVC_CHECK_ARG(flow_mode == 0 || flow_mode == 1);
    DSVC_CHECK_ARG(layout == DSVC_LAYOUT_NCHW);
    if (!grad_input && !grad_flow) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    WarpParams p{B, C, H, W, sx, sy, inv_sx, inv_sy, flow_mode};
    // 0 = auto (staged kernel when the shape is eligible), 1 = per-pixel kernel, 2 = staged forced
    const int algo = g_bwd_algo >= 3 ? 0 : g_bwd_algo;
    if (algo != 1) {
        const int r = dsvc_warp_bwd_staged_launch(grad_out, input, flow, grad_input, grad_flow, lin_x, lin_y, p,
                                                  algo == 2, st);
        if (r != -1) return r;
        if (algo == 2) return (int)cudaErrorInvalidValue;
    }
    dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8, B);
    if (grad_input && grad_flow)
        warp_bwd_nchw<true, true><<<grid, block, 0, st>>>(grad_out, input, flow, grad_input, grad_flow, lin_x, lin_y, p, nullptr);
    else if (grad_input)
        warp_bwd_nchw<true, false><<<grid, block, 0, st>>>(grad_out, input, flow, grad_input, grad_flow, lin_x, lin_y, p, nullptr);
    else
        warp_bwd_nchw<false, true><<<grid, block, 0, st>>>(grad_out, input, flow, grad_input, grad_flow, lin_x, lin_y, p, nullptr);
    DSVC_RETURN_LAST();
}

// One warp per 64 x 16 tile samples 32 of its pixels (an 8 x 4 grid) and votes whether the tile's
// source bounding box fits the staged kernel's box; the last CTA to finish turns the count into
// the launch's mode and leaves the counters zeroed.  A heuristic for speed only: both kernels are
// correct for any flow.
__global__ void __launch_bounds__(256)
warp_bwd_scout_kernel(const float* __restrict__ flow, const float* __restrict__ lin_x, const float* __restrict__ lin_y,
                      WarpParams p, int tiles_x, int tiles_y, int ntiles, int* __restrict__ state) {
    // state[0] = mode (output), state[1] = tiles that fit, state[2] = CTAs done
    const int lane = threadIdx.x & 31, tile = blockIdx.x * 8 + (threadIdx.x >> 5);
    int fits = 0;
    if (tile < ntiles) {
        const int b = tile / (tiles_x * tiles_y), r = tile - b * tiles_x * tiles_y;
        const int tx0 = (r % tiles_x) * 64, ty0 = (r / tiles_x) * 16;
        const int x = min(tx0 + (lane & 7) * 9, p.W - 1), y = min(ty0 + (lane >> 3) * 5, p.H - 1);
        const size_t plane = (size_t)p.H * p.W, pix = (size_t)y * p.W + x;
        const float* fl = flow + (size_t)b * 2 * plane;
        const BwdCoord c = bwd_coord(__ldg(lin_x + x), __ldg(lin_y + y), __ldg(fl + pix), __ldg(fl + plane + pix), p);
        const int mnx = __reduce_min_sync(0xffffffffu, c.t.x0), mxx = __reduce_max_sync(0xffffffffu, c.t.x0);
        const int mny = __reduce_min_sync(0xffffffffu, c.t.y0), mxy = __reduce_max_sync(0xffffffffu, c.t.y0);
        fits = (mxx + 1 - (mnx & ~3) + 1 <= 96 && mxy + 1 - mny + 1 <= 32) ? 1 : 0;
    }
    __shared__ int s_fit;
    if (threadIdx.x == 0) s_fit = 0;
    __syncthreads();
    if (lane == 0 && fits) atomicAdd(&s_fit, 1);
    __syncthreads();
    if (threadIdx.x == 0) {
        if (s_fit) atomicAdd(&state[1], s_fit);
        __threadfence();
        if (atomicAdd(&state[2], 1) == (int)gridDim.x - 1) {
            const int nfit = atomicAdd(&state[1], 0);
            state[0] = (2 * nfit >= ntiles) ? DSVC_BWD_MODE_STAGED : DSVC_BWD_MODE_DIRECT;
            state[1] = 0;
            __threadfence();
            state[2] = 0;
        }
    }
}

size_t dsvc_warp_bwd_gather_workspace(int B, int H, int W);  // warp_bwd_gather.cu
int dsvc_warp_bwd_gather_launch(const float* gout, const float* input, const float* flow, float* gin,
                                float* gflow, const float* lin_x, const float* lin_y, const WarpParams& p,
                                bool force, void* workspace, size_t workspace_bytes, cudaStream_t st);

size_t dsvc_warp_bwd_cell_workspace(int B, int H, int W);  // warp_bwd_cell.cu
int dsvc_warp_bwd_cell_launch(const float* gout, const float* input, const float* flow, float* gin, float* gflow,
                              const float* lin_x, const float* lin_y, const WarpParams& p, void* workspace,
                              size_t workspace_bytes, cudaStream_t st);

extern "C" size_t dsvc_warp_bwd_cell_workspace_bytes(int B, int H, int W) { return dsvc_warp_bwd_cell_workspace(B, H, W); }

extern "C" size_t dsvc_warp_bwd_workspace_bytes(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return dsvc_warp_bwd_gather_workspace(B, H, W);
}

extern "C" int dsvc_warp_bwd_ws_f32(const float* grad_out, const float* input, const float* flow,
                                    float* grad_input, float* grad_flow, int B, int C, int H, int W,
                                    const float* lin_x, const float* lin_y, float sx, float sy,
                                    float inv_sx, float inv_sy, int flow_mode, int layout,
                                    void* workspace, size_t workspace_bytes, void* stream) {
    DSVC_CHECK_ARG(warp_args_ok(grad_out, input, flow, B, C, H, W, lin_x, lin_y));
    DSVC_CHECK_ARG(flow_mode == 0 || flow_mode == 1);
    DSVC_CHECK_ARG(layout == DSVC_LAYOUT_NCHW);
    if (!grad_input && !grad_flow) return 0;
    cudaStream_t st = (cudaStream_t)stream;
    // The gather kernel is opt-in (DSVC_WARP_BWD_GATHER): measured on B200 it ties the staged scatter
    // kernel at 1080p (921 vs 943 us) and loses at 8x64x256x256 (362 vs 280 us) -- DESIGN.md 4.3
    if (grad_input && g_bwd_algo == 3) {
        WarpParams p{B, C, H, W, sx, sy, inv_sx, inv_sy, flow_mode};
        const int r = dsvc_warp_bwd_gather_launch(grad_out, input, flow, grad_input, grad_flow, lin_x, lin_y, p,
                                                  g_bwd_algo == 3, workspace, workspace_bytes, st);
        if (r != -1) return r;
        if (g_bwd_algo == 3) return (int)cudaErrorInvalidValue;
    }
    // Default for the wide warps (C >= 8) when the caller's workspace holds the cell tables
    // (dsvc_warp_bwd_cell_workspace_bytes): the cell-order kernel (csrc/warp_bwd_cell.cu) -- no zero-fill,
    // no atomics on the main path, any flow.  Measured against the scout + staged | per-pixel pair below
    // (B200, both gradients): 1080p smooth 813 vs 1 036 us, iid stress 1 254 vs 3 028, 8x64x256x256 241 vs 293.
    if (grad_input && (g_bwd_algo == 4 || (g_bwd_algo == 0 && C >= 8))) {
        WarpParams p{B, C, H, W, sx, sy, inv_sx, inv_sy, flow_mode};
        const int r = dsvc_warp_bwd_cell_launch(grad_out, input, flow, grad_input, grad_flow, lin_x, lin_y, p, workspace,
                                                workspace_bytes, st);
        if (r != -1) return r;
        if (g_bwd_algo == 4) return (int)cudaErrorInvalidValue;
    }
    if (grad_input) {
        const cudaError_t e = cudaMemsetAsync(grad_input, 0, (size_t)B * C * H * W * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
    }
    // Default choice with a workspace: a scout launch estimates how many tiles the staged kernel can
    // stage; the staged kernel and the per-pixel kernel are both enqueued and the one the scout
    // did not pick exits at once (no host synchronisation; CUDA-graph capturable).  Under wild
    // flows (iid displacements: no tile stageable) the staged kernel's in-kernel fallback is 20 %
    // slower than the per-pixel kernel it imitates; under SpyNet-like flows the pair of extra
    // launches costs ~10 us of a 0.9 ms call.
    if (g_bwd_algo == 0 && grad_input && workspace && aligned16(workspace) &&
        workspace_bytes >= dsvc_warp_bwd_gather_workspace(B, H, W)) {
        WarpParams p{B, C, H, W, sx, sy, inv_sx, inv_sy, flow_mode};
        const int tiles_x = (W + 63) / 64, tiles_y = (H + 15) / 16;
        const long long ntiles = (long long)tiles_x * tiles_y * B;
        int* state = static_cast<int*>(workspace);   // 4 ints ahead of the gather kernel's tile flags (unused here)
        if (ntiles < (1ll << 24)) {
            cudaError_t e = cudaMemsetAsync(state, 0, 16, st);
            if (e != cudaSuccess) return (int)e;
            warp_bwd_scout_kernel<<<(unsigned)((ntiles + 7) / 8), 256, 0, st>>>(flow, lin_x, lin_y, p, tiles_x, tiles_y,
                                                                                (int)ntiles, state);
            const int r = dsvc_warp_bwd_staged_launch(grad_out, input, flow, grad_input, grad_flow, lin_x, lin_y, p, false,
                                                      st, state);
            if (r == 0) {   // staged kernel enqueued: the per-pixel kernel behind it, gated the other way
                dim3 block(32, 8), grid((W + 31) / 32, (H + 7) / 8, B);
                if (grad_flow)
                    warp_bwd_nchw<true, true><<<grid, block, 0, st>>>(grad_out, input, flow, grad_input, grad_flow, lin_x, lin_y, p, state);
                else
                    warp_bwd_nchw<true, false><<<grid, block, 0, st>>>(grad_out, input, flow, grad_input, grad_flow, lin_x, lin_y, p, state);
                DSVC_RETURN_LAST();
            }
            if (r != -1) return r;
        }
    }
    return dsvc_warp_bwd_f32(grad_out, input, flow, grad_input, grad_flow, B, C, H, W, lin_x, lin_y, sx, sy, inv_sx,
                             inv_sy, flow_mode, layout, stream);
}
