// Backward of the bilinear warp: coordinate arithmetic and the per-pixel direct scatter.
//
// Restates ATen grid_sampler_2d_backward_kernel (bilinear / border / align_corners=True;
// clip_coordinates_set_grad) followed by autograd's division of the grid gradient by
// (W-1)/2, (H-1)/2 (/root/reference/modules.py:54-62 differentiated).
#pragma once
#include "warp_common.cuh"

namespace dsvc {

// Launch-level choice between the staged and the per-pixel backward kernels, made on the device
// by warp_bwd_scout_kernel (no host synchronisation: both kernels are enqueued, one of them exits).
enum : int { DSVC_BWD_MODE_STAGED = 1, DSVC_BWD_MODE_DIRECT = 2 };

struct BwdCoord {
    Taps t;
    float wx0, wx1, wy0, wy1;  // 1-D weights (ix_se - ix, ix - ix_nw, ...)
    float gx_mult, gy_mult;    // d(source coordinate)/d(grid), zero where the coordinate is clipped
};

__device__ __forceinline__ BwdCoord bwd_coord(float lx, float ly, float fx, float fy, const WarpParams& p) {
    BwdCoord r;
    // unclipped coordinate, then clip_coordinates_set_grad
    const float fsx = p.flow_mode ? __fdiv_rn(fx, p.sx) : __fmul_rn(fx, p.inv_sx);
    const float fsy = p.flow_mode ? __fdiv_rn(fy, p.sy) : __fmul_rn(fy, p.inv_sy);
    float ix = __fmul_rn(__fmul_rn(__fadd_rn(__fadd_rn(lx, fsx), 1.0f), 0.5f), (float)(p.W - 1));
    float iy = __fmul_rn(__fmul_rn(__fadd_rn(__fadd_rn(ly, fsy), 1.0f), 0.5f), (float)(p.H - 1));
    r.gx_mult = (float)(p.W - 1) * 0.5f;
    r.gy_mult = (float)(p.H - 1) * 0.5f;
    if (ix <= 0.0f) { ix = 0.0f; r.gx_mult = 0.0f; }
    else if (ix >= (float)(p.W - 1)) { ix = (float)(p.W - 1); r.gx_mult = 0.0f; }
    if (iy <= 0.0f) { iy = 0.0f; r.gy_mult = 0.0f; }
    else if (iy >= (float)(p.H - 1)) { iy = (float)(p.H - 1); r.gy_mult = 0.0f; }
    if (!(ix == ix)) { ix = 0.0f; r.gx_mult = 0.0f; }  // NaN flow: keep every address legal
    if (!(iy == iy)) { iy = 0.0f; r.gy_mult = 0.0f; }
    r.t = make_taps(ix, iy, p.W, p.H);
    const float fx0 = (float)r.t.x0, fy0 = (float)r.t.y0;
    r.wx0 = __fsub_rn(fx0 + 1.0f, ix);
    r.wx1 = __fsub_rn(ix, fx0);
    r.wy0 = __fsub_rn(fy0 + 1.0f, iy);
    r.wy1 = __fsub_rn(iy, fy0);
    return r;
}

// grad_flow of one pixel from the channel-summed grid gradient (autograd of flow / s: grad / s;
// ATen's CUDA branch multiplies by the fp32 reciprocal).
__device__ __forceinline__ void store_gflow(float* __restrict__ gflow, const WarpParams& p, int b,
                                            size_t pix, const BwdCoord& bc, float gix, float giy,
                                            bool accumulate) {
    const size_t plane = (size_t)p.H * p.W;
    float* gf = gflow + (size_t)b * 2 * plane + pix;
    const float ggx = __fmul_rn(bc.gx_mult, gix), ggy = __fmul_rn(bc.gy_mult, giy);
    const float vx = p.flow_mode ? __fdiv_rn(ggx, p.sx) : __fmul_rn(ggx, p.inv_sx);
    const float vy = p.flow_mode ? __fdiv_rn(ggy, p.sy) : __fmul_rn(ggy, p.inv_sy);
    if (accumulate) {  // channel ranges of one pixel handled by several CTAs (gflow pre-zeroed)
        atomicAdd(gf, vx);
        atomicAdd(gf + plane, vy);
    } else {
        gf[0] = vx;
        gf[plane] = vy;
    }
}

// One thread: output pixel (x, y) of batch item b, channels [c0, c1).  grad_flow is a
// per-pixel reduction over the channels kept in registers; grad_input taps are scattered
// with float reductions (RED.ADD.F32, no return value).
template <bool NEED_GIN, bool NEED_GFLOW>
__device__ __forceinline__ void bwd_pixel_direct(const float* __restrict__ gout, const float* __restrict__ in,
                                                 const float* __restrict__ flow, float* __restrict__ gin,
                                                 float* __restrict__ gflow, const float* __restrict__ lin_x,
                                                 const float* __restrict__ lin_y, const WarpParams& p,
                                                 int b, int x, int y, int c0, int c1, bool accumulate_gflow) {
    const size_t plane = (size_t)p.H * p.W;
    const size_t pix = (size_t)y * p.W + x;
    const float* fl = flow + (size_t)b * 2 * plane + pix;
    const BwdCoord bc = bwd_coord(__ldg(lin_x + x), __ldg(lin_y + y), __ldg(fl), __ldg(fl + plane), p);
    const Taps& t = bc.t;
    const float wx0 = bc.wx0, wx1 = bc.wx1, wy0 = bc.wy0, wy1 = bc.wy1;
    const int o_nw = t.y0 * p.W + t.x0;
    const int dx = t.x1ok ? 1 : 0, dy = t.y1ok ? p.W : 0;
    const bool xe = t.x1ok, ys = t.y1ok, xy = t.x1ok && t.y1ok;
    const float* gp = gout + ((size_t)b * p.C + c0) * plane + pix;
    const float* ip = in + ((size_t)b * p.C + c0) * plane + o_nw;
    float* gi = NEED_GIN ? gin + ((size_t)b * p.C + c0) * plane + o_nw : nullptr;
    float gix = 0.0f, giy = 0.0f;
#pragma unroll 4
    for (int c = c0; c < c1; ++c) {
        const float g = __ldg(gp);
        if (NEED_GIN) {
            atomicAdd(gi, __fmul_rn(t.nw, g));
            if (xe) atomicAdd(gi + dx, __fmul_rn(t.ne, g));
            if (ys) atomicAdd(gi + dy, __fmul_rn(t.sw, g));
            if (xy) atomicAdd(gi + dy + dx, __fmul_rn(t.se, g));
            gi += plane;
        }
        if (NEED_GFLOW) {
            const float v_nw = __ldg(ip);
            gix -= v_nw * wy0 * g;
            giy -= v_nw * wx0 * g;
            if (xe) {
                const float v = __ldg(ip + dx);
                gix += v * wy0 * g;
                giy -= v * wx1 * g;
            }
            if (ys) {
                const float v = __ldg(ip + dy);
                gix -= v * wy1 * g;
                giy += v * wx0 * g;
            }
            if (xy) {
                const float v = __ldg(ip + dy + dx);
                gix += v * wy1 * g;
                giy += v * wx1 * g;
            }
            ip += plane;
        }
        gp += plane;
    }
    if (NEED_GFLOW) store_gflow(gflow, p, b, pix, bc, gix, giy, accumulate_gflow);
}

}  // namespace dsvc
