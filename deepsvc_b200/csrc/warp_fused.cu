// Fusions on either side of the few-channel warps, NCHW fp32, sm_100a.
//
//  * SpyNet level (/root/reference/modules.py:163-168): flow_up = bilinearupsacling(flow) * 2.0
//    (F.interpolate x2, bilinear, align_corners=False, modules.py:107-112) followed by
//    torch_warp(im2, flow_up).  One launch reads the coarse flow, writes flow_up (the level's
//    network input and residual base) and the warped image: 34 B/pixel instead of 58.
//  * Frame warp + warp_loss (/root/reference/video_model.py:37-38):
//    warped = torch_warp(ref, mv); warp_loss = mean((warped - cur)^2).  The squared error is
//    reduced per CTA in a fixed order (double partials, finalised by dsvc_bits_finalize_f64
//    with scale 1 / numel): 12 B/pixel extra instead of three more elementwise passes.
//
// The upsampling restates ATen's upsample_bilinear2d (UpSampleBilinear2d.cu: source index
// 0.5 * (dst + 0.5) - 0.5 clamped at 0, weights 1 - lambda / lambda, rows blended after
// columns); the weights are exact binary fractions, products may be contracted differently
// from ATen's build, so flow_up agrees to 1 ulp and the warped image to the usual 1e-5.
#include <algorithm>

#include "warp_common.cuh"
#include "../../include/deepsvc_b200.h"

namespace dsvc {

constexpr int kFusedThreads = 256;

__device__ __forceinline__ float upsample2_tap(const float* __restrict__ src, int h2, int w2, int x, int y) {
    const float sy = fmaxf(__fmaf_rn(0.5f, (float)y + 0.5f, -0.5f), 0.0f);  // exact in fp32
    const float sx = fmaxf(__fmaf_rn(0.5f, (float)x + 0.5f, -0.5f), 0.0f);
    const int y1 = (int)sy, x1 = (int)sx;
    const int yp = y1 < h2 - 1 ? w2 : 0, xp = x1 < w2 - 1 ? 1 : 0;
    const float ly1 = sy - (float)y1, ly0 = 1.0f - ly1;
    const float lx1 = sx - (float)x1, lx0 = 1.0f - lx1;
    const float* q = src + (size_t)y1 * w2 + x1;
    const float a = __ldg(q), b = __ldg(q + xp), c = __ldg(q + yp), d = __ldg(q + yp + xp);
    return ly0 * (lx0 * a + lx1 * b) + ly1 * (lx0 * c + lx1 * d);
}

template <int C, bool UPS, bool MSE>
__global__ void __launch_bounds__(kFusedThreads)
warp_fused_kernel(const float* __restrict__ in, const float* __restrict__ flow,
                  const float* __restrict__ flow_coarse, float* __restrict__ flow_up,
                  const float* __restrict__ target, double* __restrict__ sq_partials,
                  float* __restrict__ out, const float* __restrict__ lin_x,
                  const float* __restrict__ lin_y, WarpParams p) {
    __shared__ double s_red[kFusedThreads / 32];
    const int x = blockIdx.x * 32 + (threadIdx.x & 31), y = blockIdx.y * 8 + (threadIdx.x >> 5);
    const int b = blockIdx.z;
    const size_t plane = (size_t)p.H * p.W;
    float sq = 0.0f;
    if (x < p.W && y < p.H) {
        const size_t pix = (size_t)y * p.W + x;
        float fx, fy;
        if (UPS) {
            const int h2 = p.H >> 1, w2 = p.W >> 1;
            const float* fc = flow_coarse + (size_t)b * 2 * h2 * w2;
            fx = upsample2_tap(fc, h2, w2, x, y) * 2.0f;  // modules.py:163
            fy = upsample2_tap(fc + (size_t)h2 * w2, h2, w2, x, y) * 2.0f;
            float* fu = flow_up + (size_t)b * 2 * plane + pix;
            fu[0] = fx;
            fu[plane] = fy;
        } else {
            const float* fl = flow + (size_t)b * 2 * plane + pix;
            fx = __ldg(fl);
            fy = __ldg(fl + plane);
        }
        const float ix = source_coord(__ldg(lin_x + x), fx, p.sx, p.inv_sx, p.flow_mode, p.W);
        const float iy = source_coord(__ldg(lin_y + y), fy, p.sy, p.inv_sy, p.flow_mode, p.H);
        const Taps t = make_taps(ix, iy, p.W, p.H);
        const int dx = t.x1ok ? 1 : 0, dy = t.y1ok ? p.W : 0;
        const float* ip = in + (size_t)b * C * plane + (size_t)t.y0 * p.W + t.x0;
        float* op = out + (size_t)b * C * plane + pix;
        const float* tp = MSE ? target + (size_t)b * C * plane + pix : nullptr;
        float v[C][4], tv[C];
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const float* q = ip + (size_t)c * plane;
            v[c][0] = __ldg(q); v[c][1] = __ldg(q + dx); v[c][2] = __ldg(q + dy); v[c][3] = __ldg(q + dy + dx);
            if (MSE) tv[c] = __ldg(tp + (size_t)c * plane);
        }
#pragma unroll
        for (int c = 0; c < C; ++c) {
            float acc = __fmul_rn(v[c][0], t.nw);  // same order of operations as gather_pixel
            acc = t.x1ok ? fmaf(v[c][1], t.ne, acc) : acc;
            acc = t.y1ok ? fmaf(v[c][2], t.sw, acc) : acc;
            acc = (t.x1ok && t.y1ok) ? fmaf(v[c][3], t.se, acc) : acc;
            op[(size_t)c * plane] = acc;
            if (MSE) {
                const float d = acc - tv[c];
                sq = fmaf(d, d, sq);
            }
        }
    }
    if (MSE) {
        const double r = block_sum((double)sq, s_red);
        if (threadIdx.x == 0)
            sq_partials[((size_t)blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x] = r;
    }
}

}  // namespace dsvc

using namespace dsvc;

extern "C" int dsvc_warp_fused_slots(int B, int H, int W) {
    if (B <= 0 || H <= 0 || W <= 0) return 0;
    return B * ((H + 7) / 8) * ((W + 31) / 32);
}

extern "C" int dsvc_warp_fused_f32(const float* input, const float* flow, const float* flow_coarse,
                                   float* flow_up, const float* target, double* sq_partials,
                                   float* out, int B, int C, int H, int W, const float* lin_x,
                                   const float* lin_y, float sx, float sy, float inv_sx,
                                   float inv_sy, int flow_mode, void* stream) {
    DSVC_CHECK_ARG(input && out && lin_x && lin_y && B > 0 && C > 0 && C <= 4 && H > 1 && W > 1);
    DSVC_CHECK_ARG((int64_t)H * W < (int64_t)1 << 30 && B <= 65535);
    DSVC_CHECK_ARG(flow_mode == 0 || flow_mode == 1);
    const bool ups = flow_coarse != nullptr, mse = target != nullptr;
    DSVC_CHECK_ARG(ups != (flow != nullptr));                 // exactly one flow source
    DSVC_CHECK_ARG(!ups || (flow_up && H % 2 == 0 && W % 2 == 0));
    DSVC_CHECK_ARG(mse == (sq_partials != nullptr));
    WarpParams p{B, C, H, W, sx, sy, inv_sx, inv_sy, flow_mode};
    dim3 grid((W + 31) / 32, (H + 7) / 8, B);
    cudaStream_t st = (cudaStream_t)stream;
#define DSVC_FUSED(CN)                                                                                        \
    if (ups && mse) warp_fused_kernel<CN, true, true><<<grid, kFusedThreads, 0, st>>>(input, flow, flow_coarse, flow_up, target, sq_partials, out, lin_x, lin_y, p); \
    else if (ups) warp_fused_kernel<CN, true, false><<<grid, kFusedThreads, 0, st>>>(input, flow, flow_coarse, flow_up, target, sq_partials, out, lin_x, lin_y, p); \
    else if (mse) warp_fused_kernel<CN, false, true><<<grid, kFusedThreads, 0, st>>>(input, flow, flow_coarse, flow_up, target, sq_partials, out, lin_x, lin_y, p); \
    else warp_fused_kernel<CN, false, false><<<grid, kFusedThreads, 0, st>>>(input, flow, flow_coarse, flow_up, target, sq_partials, out, lin_x, lin_y, p); \
    break
    switch (C) {
        case 1: DSVC_FUSED(1);
        case 2: DSVC_FUSED(2);
        case 3: DSVC_FUSED(3);
        default: DSVC_FUSED(4);
    }
#undef DSVC_FUSED
    DSVC_RETURN_LAST();
}

// ------------------------------------------------------------------ blend (modules.py:436)
// out = w * warped + (1 - w) * pred, elementwise on [B,3,H,W]: one pass (16 B/element) instead
// of the reference's four elementwise launches.  Same operation order as the reference's
// expression (two products, one subtraction, one addition; no FMA contraction).
namespace dsvc {
__global__ void __launch_bounds__(256)
blend_kernel(const float4* __restrict__ w, const float4* __restrict__ a, const float4* __restrict__ b,
             float4* __restrict__ out, size_t n4, const float* __restrict__ ws, const float* __restrict__ as,
             const float* __restrict__ bs, float* __restrict__ os, size_t n) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    auto f = [](float wv, float av, float bv) {
        return __fadd_rn(__fmul_rn(wv, av), __fmul_rn(__fsub_rn(1.0f, wv), bv));
    };
    if (i < n4) {
        const float4 wv = __ldg(w + i), av = __ldg(a + i), bv = __ldg(b + i);
        out[i] = make_float4(f(wv.x, av.x, bv.x), f(wv.y, av.y, bv.y), f(wv.z, av.z, bv.z), f(wv.w, av.w, bv.w));
    }
    if (i < n - 4 * n4) {  // scalar tail
        const size_t j = 4 * n4 + i;
        os[j] = f(__ldg(ws + j), __ldg(as + j), __ldg(bs + j));
    }
}
}  // namespace dsvc

extern "C" int dsvc_blend_f32(const float* weight, const float* warped, const float* pred, float* out,
                              int64_t n, void* stream) {
    DSVC_CHECK_ARG(weight && warped && pred && out && n >= 0);
    if (n == 0) return 0;
    const bool vec = aligned16(weight) && aligned16(warped) && aligned16(pred) && aligned16(out);
    const size_t n4 = vec ? (size_t)n / 4 : 0;
    const size_t threads = std::max(n4, (size_t)n - 4 * n4);
    blend_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        reinterpret_cast<const float4*>(weight), reinterpret_cast<const float4*>(warped),
        reinterpret_cast<const float4*>(pred), reinterpret_cast<float4*>(out), n4, weight, warped, pred, out, (size_t)n);
    DSVC_RETURN_LAST();
}

// ------------------------------------------------------------------ LRP add (image_model.py:185-188)
// `lrp = 0.5 * torch.tanh(lrp); y_hat_slice += lrp`: three eager launches on a [B,Cs,H/16,W/16]
// tensor per slice (16 slices per frame) -> one.  Same fp32 operations in the same order
// (tanhf, one multiply, one add): bit-identical.  Differentiable form: d/d y_hat = 1,
// d/d lrp = 0.5 * (1 - tanh^2), written to grad_lrp when asked for.
namespace dsvc {
__global__ void __launch_bounds__(256)
lrp_add_kernel(const float* __restrict__ y_hat, const float* __restrict__ lrp, float* __restrict__ out, size_t n) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) out[i] = __fadd_rn(y_hat[i], __fmul_rn(0.5f, tanhf(lrp[i])));
}
__global__ void __launch_bounds__(256)
lrp_add_bwd_kernel(const float* __restrict__ grad_out, const float* __restrict__ lrp, float* __restrict__ grad_lrp, size_t n) {
    const size_t i = (size_t)blockIdx.x * 256 + threadIdx.x;
    if (i < n) {
        const float t = tanhf(lrp[i]);
        grad_lrp[i] = grad_out[i] * (0.5f * (1.0f - t * t));
    }
}
}  // namespace dsvc

extern "C" int dsvc_lrp_add_f32(const float* y_hat, const float* lrp, float* out, int64_t n, void* stream) {
    DSVC_CHECK_ARG(y_hat && lrp && out && n >= 0);
    if (n == 0) return 0;
    dsvc::lrp_add_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(y_hat, lrp, out, (size_t)n);
    DSVC_RETURN_LAST();
}

extern "C" int dsvc_lrp_add_bwd_f32(const float* grad_out, const float* lrp, float* grad_lrp, int64_t n, void* stream) {
    DSVC_CHECK_ARG(grad_out && lrp && grad_lrp && n >= 0);
    if (n == 0) return 0;
    dsvc::lrp_add_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(grad_out, lrp, grad_lrp, (size_t)n);
    DSVC_RETURN_LAST();
}

