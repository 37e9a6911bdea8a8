// Coordinate arithmetic of the reference's backward warp, restated op for op.
//
// Reference: /root/reference/modules.py:25-62 (torch_warp) -> F.grid_sample(bilinear,
// border, align_corners=True) -> ATen grid_sampler_2d (GridSampler.cuh:21-57
// grid_sampler_unnormalize / clip_coordinates, GridSampler.cu forward/backward
// kernels).  Every fp32 rounding step of that chain is kept (explicit _rn
// intrinsics forbid FMA contraction across normalise / add / unnormalise), because
// a "direct" x + flow formulation differs by up to 2e-4 at 1080p.
#pragma once
#include "common.cuh"

namespace dsvc {

struct WarpParams {
    int B, C, H, W;
    float sx, sy;          // (W-1)/2, (H-1)/2 as fp32
    float inv_sx, inv_sy;  // fp32 reciprocals as ATen's CUDA div-by-scalar computes them
    int flow_mode;         // DSVC_FLOW_MUL_RECIPROCAL / DSVC_FLOW_TRUE_DIVIDE
};

struct Taps {
    int x0, y0;            // north-west tap (always inside the image)
    bool x1ok, y1ok;       // east / south taps inside the image
    float nw, ne, sw, se;  // bilinear weights, ATen order of operations
};

// normalised grid value -> clamped source coordinate (align_corners=True, border)
__device__ __forceinline__ float source_coord(float lin, float fl, float s, float inv_s,
                                              int flow_mode, int size) {
    const float fs = flow_mode ? __fdiv_rn(fl, s) : __fmul_rn(fl, inv_s);  // modules.py:54-55
    const float g = __fadd_rn(lin, fs);                                    // modules.py:57
    // grid_sampler_unnormalize: ((g + 1) / 2) * (size - 1)
    float c = __fmul_rn(__fmul_rn(__fadd_rn(g, 1.0f), 0.5f), (float)(size - 1));
    // clip_coordinates: min(size - 1, max(c, 0))
    c = fminf((float)(size - 1), fmaxf(c, 0.0f));
    return c;
}

__device__ __forceinline__ Taps make_taps(float ix, float iy, int W, int H) {
    Taps t;
    const float fx0 = floorf(ix), fy0 = floorf(iy);
    t.x0 = (int)fx0;
    t.y0 = (int)fy0;
    t.x1ok = t.x0 + 1 < W;
    t.y1ok = t.y0 + 1 < H;
    const float wx0 = __fsub_rn(fx0 + 1.0f, ix);  // ix_se - ix
    const float wx1 = __fsub_rn(ix, fx0);         // ix - ix_nw
    const float wy0 = __fsub_rn(fy0 + 1.0f, iy);
    const float wy1 = __fsub_rn(iy, fy0);
    t.nw = __fmul_rn(wx0, wy0);
    t.ne = __fmul_rn(wx1, wy0);
    t.sw = __fmul_rn(wx0, wy1);
    t.se = __fmul_rn(wx1, wy1);
    return t;
}

// One thread: output pixel (x, y) of batch item b, channels [c0, cend) -- the direct
// read-only-path gather (the gather kernel, and the staged kernel for rectangles it cannot stage).
template <int UNROLL>
__device__ __forceinline__ void gather_pixel(const float* __restrict__ in,
                                             const float* __restrict__ flow,
                                             float* __restrict__ out,
                                             const float* __restrict__ lin_x,
                                             const float* __restrict__ lin_y, const WarpParams& p,
                                             int b, int x, int y, int c0, int cend) {
    const size_t plane = (size_t)p.H * p.W;
    const size_t pix = (size_t)y * p.W + x;
    const float* fl = flow + (size_t)b * 2 * plane + pix;
    const float fx = __ldg(fl), fy = __ldg(fl + plane);
    const float ix = source_coord(__ldg(lin_x + x), fx, p.sx, p.inv_sx, p.flow_mode, p.W);
    const float iy = source_coord(__ldg(lin_y + y), fy, p.sy, p.inv_sy, p.flow_mode, p.H);
    const Taps t = make_taps(ix, iy, p.W, p.H);
    const int o_nw = t.y0 * p.W + t.x0;
    const int dx = t.x1ok ? 1 : 0;    // clamped so that the address stays legal;
    const int dy = t.y1ok ? p.W : 0;  // the value is discarded when the tap is outside
    const float* ip = in + ((size_t)b * p.C + c0) * plane + o_nw;
    float* op = out + ((size_t)b * p.C + c0) * plane + pix;
    int c = c0;
    for (; c + UNROLL <= cend; c += UNROLL) {
        float v[UNROLL][4];
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            const float* q = ip + (size_t)u * plane;
            v[u][0] = __ldg(q);
            v[u][1] = __ldg(q + dx);
            v[u][2] = __ldg(q + dy);
            v[u][3] = __ldg(q + dy + dx);
        }
#pragma unroll
        for (int u = 0; u < UNROLL; ++u) {
            float acc = __fmul_rn(v[u][0], t.nw);
            acc = t.x1ok ? fmaf(v[u][1], t.ne, acc) : acc;
            acc = t.y1ok ? fmaf(v[u][2], t.sw, acc) : acc;
            acc = (t.x1ok && t.y1ok) ? fmaf(v[u][3], t.se, acc) : acc;
            st_stream1(op + (size_t)u * plane, acc);
        }
        ip += (size_t)UNROLL * plane;
        op += (size_t)UNROLL * plane;
    }
    for (; c < cend; ++c) {
        float acc = __fmul_rn(__ldg(ip), t.nw);
        acc = t.x1ok ? fmaf(__ldg(ip + dx), t.ne, acc) : acc;
        acc = t.y1ok ? fmaf(__ldg(ip + dy), t.sw, acc) : acc;
        acc = (t.x1ok && t.y1ok) ? fmaf(__ldg(ip + dy + dx), t.se, acc) : acc;
        st_stream1(op, acc);
        ip += plane;
        op += plane;
    }
}

// Scheduler state of the persistent staged kernel (warp_persist.cu).  Lives in the caller's
// workspace, which is all-zero before and after every launch.
struct WarpSched {
    int next;    // work units claimed so far
    int exited;  // CTAs that have run dry; the last one re-zeroes the state
    int pad[2];
};

}  // namespace dsvc
