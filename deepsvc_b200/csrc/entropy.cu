// Hyperprior entropy model: fused quantise / likelihood / bit-estimate kernels.
//
// Replaces compressai 1.2.1 GaussianConditional / EntropyBottleneck / ste_round /
// LowerBound as called from /root/reference/image_model.py:155,160-162,181,183,
// 237-238,286-290 and the bit sums of video_model.py:39-42,53-56.  The eager
// reference issues ~18 elementwise launches per GaussianConditional.forward, ~190
// per build_indexes and ~70 per EntropyBottleneck.forward; here each call is ONE
// launch that reads every input once and writes every requested output once
// (16 B/element for the eval path with fused bits).  Elementwise + reduction work:
// HBM/latency bound, no tensor cores.
//
// Each fp32 rounding of the reference chain is kept (explicit _rn intrinsics, true
// division, erfcf / tanhf / expf / logf of the CUDA math library -- the same
// functions ATen's CUDA kernels call), so quantised symbols and table indexes are
// bit-exact and likelihoods agree with stock torch to the last bits.
#include "common.cuh"
#include "../../include/deepsvc_b200.h"

namespace dsvc {

constexpr int kGcThreads = 128;
constexpr int kGcVec = 4;
constexpr int kMaxTable = 256;

struct GcArgs {
    const float* x;
    const float* scales;
    const float* means;
    const float* noise;
    float* outputs;
    float* likelihood;
    float* y_hat;
    int32_t* symbols;
    int32_t* indexes;
    const float* scale_table;
    int n_table;
    double* bits_partials;
    float scale_bound, lik_bound;
    long long inner, x_rs, scales_rs, means_rs, noise_rs;
};

__device__ __forceinline__ float lower_bound_nan(float v, float bound) {
    // torch.max propagates NaN
    return (v != v) ? v : fmaxf(v, bound);
}

__device__ __forceinline__ float std_cumulative(float t) {
    // GaussianConditional._standardized_cumulative: 0.5 * erfc(-(2**-0.5) * t)
    const float kConst = -0.70710678118654752440f;
    return __fmul_rn(0.5f, erfcf(__fmul_rn(kConst, t)));
}

struct GcElem {
    float out, lik, yhat, q, s;
};

__device__ __forceinline__ GcElem gc_element(float x, float sc, float mu, float nz, bool noisy,
                                             float scale_bound, float lik_bound) {
    GcElem e;
    e.q = rintf(__fsub_rn(x, mu));         // torch.round: half to even
    e.yhat = __fadd_rn(e.q, mu);           // ste_round(y - mu) + mu  ==  quantize("dequantize")
    e.out = noisy ? __fadd_rn(x, nz) : e.yhat;
    const float v = fabsf(__fsub_rn(e.out, mu));  // the reference re-subtracts the mean
    e.s = lower_bound_nan(sc, scale_bound);
    const float up = std_cumulative(__fdiv_rn(__fsub_rn(0.5f, v), e.s));
    const float lo = std_cumulative(__fdiv_rn(__fsub_rn(-0.5f, v), e.s));
    e.lik = lower_bound_nan(__fsub_rn(up, lo), lik_bound);
    return e;
}

__device__ __forceinline__ int table_index(const float* tbl, int n_table, float s) {
    // #{k < n_table-1 : tbl[k] < s}  ==  (n_table-1) - #{k < n_table-1 : s <= tbl[k]}
    int lo = 0, hi = n_table - 1;
    while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (tbl[mid] < s) lo = mid + 1; else hi = mid;
    }
    return lo;
}

// Every thread owns kGcVec elements: one 128-bit access per tensor on the vector
// path, kGcVec coalesced 32-bit accesses (stride kGcThreads) on the scalar path, so
// the grid -- and the number of bit partials -- does not depend on alignment.
template <bool VEC>
__global__ void __launch_bounds__(kGcThreads) gc_fwd_kernel(GcArgs a) {
    __shared__ float s_tbl[kMaxTable];
    __shared__ double s_red[kGcThreads / 32];
    pdl_prologue();
    const bool want_idx = a.indexes != nullptr;
    if (want_idx) {
        for (int i = threadIdx.x; i < a.n_table; i += blockDim.x) s_tbl[i] = a.scale_table[i];
        __syncthreads();
    }
    const long long row = blockIdx.y;
    const long long base = (long long)blockIdx.x * (kGcThreads * kGcVec);
    const bool noisy = a.noise != nullptr;
    const float* px = a.x + row * a.x_rs;
    const float* ps = a.scales + row * a.scales_rs;
    const float* pm = a.means ? a.means + row * a.means_rs : nullptr;
    const float* pn = noisy ? a.noise + row * a.noise_rs : nullptr;
    const long long orow = row * a.inner;  // dense outputs
    float acc = 0.0f;
    float xv[kGcVec], sv[kGcVec], mv[kGcVec], nv[kGcVec];
    long long idx[kGcVec];
    bool ok[kGcVec];
#pragma unroll
    for (int k = 0; k < kGcVec; ++k) {
        idx[k] = VEC ? base + (long long)threadIdx.x * kGcVec + k
                     : base + (long long)k * kGcThreads + threadIdx.x;
        ok[k] = idx[k] < a.inner;
        mv[k] = 0.0f;
        nv[k] = 0.0f;
    }
    if (VEC) {
        if (ok[0]) {  // inner % 4 == 0: all four or none
            const float4 t0 = ld_stream4(px + idx[0]);
            const float4 t1 = ld_stream4(ps + idx[0]);
            xv[0] = t0.x; xv[1] = t0.y; xv[2] = t0.z; xv[3] = t0.w;
            sv[0] = t1.x; sv[1] = t1.y; sv[2] = t1.z; sv[3] = t1.w;
            if (pm) {
                const float4 t2 = ld_stream4(pm + idx[0]);
                mv[0] = t2.x; mv[1] = t2.y; mv[2] = t2.z; mv[3] = t2.w;
            }
            if (pn) {
                const float4 t3 = ld_stream4(pn + idx[0]);
                nv[0] = t3.x; nv[1] = t3.y; nv[2] = t3.z; nv[3] = t3.w;
            }
        }
    } else {
#pragma unroll
        for (int k = 0; k < kGcVec; ++k) {
            if (ok[k]) {
                xv[k] = __ldg(px + idx[k]);
                sv[k] = __ldg(ps + idx[k]);
                if (pm) mv[k] = __ldg(pm + idx[k]);
                if (pn) nv[k] = __ldg(pn + idx[k]);
            }
        }
    }
    float ov[kGcVec], lv[kGcVec], yv[kGcVec];
    int qv[kGcVec], iv[kGcVec];
#pragma unroll
    for (int k = 0; k < kGcVec; ++k) {
        if (ok[k]) {
            const GcElem e = gc_element(xv[k], sv[k], mv[k], nv[k], noisy, a.scale_bound, a.lik_bound);
            ov[k] = e.out; lv[k] = e.lik; yv[k] = e.yhat;
            qv[k] = __float2int_rz(e.q);
            iv[k] = want_idx ? table_index(s_tbl, a.n_table, e.s) : 0;
            if (a.bits_partials) acc += logf(e.lik);
        }
    }
    if (VEC) {
        if (ok[0]) {
            const long long o = orow + idx[0];
            if (a.outputs) st_stream4(a.outputs + o, make_float4(ov[0], ov[1], ov[2], ov[3]));
            if (a.likelihood) st_stream4(a.likelihood + o, make_float4(lv[0], lv[1], lv[2], lv[3]));
            if (a.y_hat) st_stream4(a.y_hat + o, make_float4(yv[0], yv[1], yv[2], yv[3]));
            if (a.symbols) *reinterpret_cast<int4*>(a.symbols + o) = make_int4(qv[0], qv[1], qv[2], qv[3]);
            if (a.indexes) *reinterpret_cast<int4*>(a.indexes + o) = make_int4(iv[0], iv[1], iv[2], iv[3]);
        }
    } else {
#pragma unroll
        for (int k = 0; k < kGcVec; ++k) {
            if (ok[k]) {
                const long long o = orow + idx[k];
                if (a.outputs) a.outputs[o] = ov[k];
                if (a.likelihood) a.likelihood[o] = lv[k];
                if (a.y_hat) a.y_hat[o] = yv[k];
                if (a.symbols) a.symbols[o] = qv[k];
                if (a.indexes) a.indexes[o] = iv[k];
            }
        }
    }
    if (a.bits_partials) {
        const double r = block_sum((double)acc, s_red);
        if (threadIdx.x == 0) a.bits_partials[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = r;
    }
}

// ------------------------------------------------------------------ GC backward
struct GcBwdArgs {
    const float* grad_lik;
    const float* x;
    const float* scales;
    const float* means;
    const float* noise;
    float* grad_x;
    float* grad_scales;
    float* grad_means;
    float scale_bound, lik_bound;
    long long inner, x_rs, scales_rs, means_rs, noise_rs;
};

__global__ void __launch_bounds__(kGcThreads) gc_bwd_kernel(GcBwdArgs a) {
    const long long row = blockIdx.y;
    const long long i = (long long)blockIdx.x * kGcThreads + threadIdx.x;
    if (i >= a.inner) return;
    const long long o = row * a.inner + i;
    const bool noisy = a.noise != nullptr;
    const float x = __ldg(a.x + row * a.x_rs + i);
    const float sc = __ldg(a.scales + row * a.scales_rs + i);
    const float mu = a.means ? __ldg(a.means + row * a.means_rs + i) : 0.0f;
    const float nz = noisy ? __ldg(a.noise + row * a.noise_rs + i) : 0.0f;
    float g = __ldg(a.grad_lik + o);
    // recompute forward
    const float q = rintf(x - mu);
    const float out = noisy ? x + nz : q + mu;
    const float dlt = out - mu;
    const float v = fabsf(dlt);
    const float s = lower_bound_nan(sc, a.scale_bound);
    const float ta = (0.5f - v) / s, tb = (-0.5f - v) / s;
    const float raw = std_cumulative(ta) - std_cumulative(tb);
    // LowerBound backward: pass if (x >= bound) | (grad < 0)
    if (!(raw >= a.lik_bound || g < 0.0f)) g = 0.0f;
    // d/dt [0.5 erfc(-t/sqrt2)] = exp(-t^2/2) / sqrt(2 pi)
    const float kInvSqrt2Pi = 0.39894228040143267794f;
    const float pa = kInvSqrt2Pi * expf(-0.5f * ta * ta);
    const float pb = kInvSqrt2Pi * expf(-0.5f * tb * tb);
    const float inv_s = 1.0f / s;
    const float dL_dv = (pb - pa) * inv_s;
    const float dL_ds = (tb * pb - ta * pa) * inv_s;
    float gs = g * dL_ds;
    if (!(sc >= a.scale_bound || gs < 0.0f)) gs = 0.0f;
    const float sgn = dlt > 0.0f ? 1.0f : (dlt < 0.0f ? -1.0f : 0.0f);
    // noise mode: o = x + n  -> dv/dx = sgn, dv/dmu = -sgn; round mode: both vanish
    const float gx = noisy ? g * dL_dv * sgn : 0.0f;
    if (a.grad_x) a.grad_x[o] = gx;
    if (a.grad_means) a.grad_means[o] = -gx;
    if (a.grad_scales) a.grad_scales[o] = gs;
}

// ------------------------------------------------------------------ EntropyBottleneck
// packed per-channel parameter layout (DSVC_EB_PARAMS_PER_CHANNEL floats):
//   [0..2]  softplus(M0) (3x1)   [3..5]  b0   [6..8]  tanh(f0)
//   k=1..3 at base 9+15(k-1): softplus(Mk) (3x3, row major [out][in]), bk (3), tanh(fk) (3)
//   [54..56] softplus(M4) (1x3)  [57] b4      [58] median   [59] pad
constexpr int kEbThreads = 128;
constexpr int kEbP = DSVC_EB_PARAMS_PER_CHANNEL;

struct EbChain {
    float h[4][3];  // inputs of layers 1..4 (h[0] = output of layer 0, ...)
    float t[4][3];  // tanh(pre-activation) of layers 0..3
    float out;
};

template <bool SAVE>
__device__ __forceinline__ float eb_logits(const float* P, float x, EbChain* ch) {
    float h[3];
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        float a = __fadd_rn(__fmul_rn(P[j], x), P[3 + j]);
        const float t = tanhf(a);
        if (SAVE) ch->t[0][j] = t;
        h[j] = __fadd_rn(a, __fmul_rn(P[6 + j], t));
        if (SAVE) ch->h[0][j] = h[j];
    }
#pragma unroll
    for (int k = 1; k <= 3; ++k) {
        const float* M = P + 9 + 15 * (k - 1);
        float g[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            float a = __fmul_rn(M[3 * j], h[0]);
            a = fmaf(M[3 * j + 1], h[1], a);
            a = fmaf(M[3 * j + 2], h[2], a);
            a = __fadd_rn(a, M[9 + j]);
            const float t = tanhf(a);
            if (SAVE) ch->t[k][j] = t;
            g[j] = __fadd_rn(a, __fmul_rn(M[12 + j], t));
        }
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            h[j] = g[j];
            if (SAVE) ch->h[k][j] = g[j];
        }
    }
    float o = __fmul_rn(P[54], h[0]);
    o = fmaf(P[55], h[1], o);
    o = fmaf(P[56], h[2], o);
    o = __fadd_rn(o, P[57]);
    if (SAVE) ch->out = o;
    return o;
}

__device__ __forceinline__ float sigmoidf_ref(float x) {
    return __fdiv_rn(1.0f, __fadd_rn(1.0f, expf(-x)));
}

__device__ __forceinline__ float signf(float x) {
    return x > 0.0f ? 1.0f : (x < 0.0f ? -1.0f : 0.0f);
}

__global__ void __launch_bounds__(kEbThreads)
eb_fwd_kernel(const float* __restrict__ z, const float* __restrict__ noise,
              const float* __restrict__ params, float* __restrict__ outputs,
              float* __restrict__ likelihood, float* __restrict__ z_hat,
              double* __restrict__ bits_partials, float lik_bound, int B, int C, int S) {
    __shared__ float P[kEbP];
    __shared__ double s_red[kEbThreads / 32];
    pdl_prologue();
    const int c = blockIdx.x;
    for (int i = threadIdx.x; i < kEbP; i += blockDim.x) P[i] = params[(size_t)c * kEbP + i];
    __syncthreads();
    const int e = blockIdx.y * kEbThreads + threadIdx.x;  // index into B*S
    float acc = 0.0f;
    if (e < B * S) {
        const int b = e / S, s = e - b * S;
        const size_t idx = ((size_t)b * C + c) * S + s;
        const float zv = __ldg(z + idx);
        const float med = P[58];
        const float zh = __fadd_rn(rintf(__fsub_rn(zv, med)), med);
        const float o = noise ? __fadd_rn(zv, __ldg(noise + idx)) : zh;
        const float lower = eb_logits<false>(P, __fsub_rn(o, 0.5f), nullptr);
        const float upper = eb_logits<false>(P, __fadd_rn(o, 0.5f), nullptr);
        const float sg = -signf(__fadd_rn(lower, upper));
        const float lik_raw = fabsf(__fsub_rn(sigmoidf_ref(__fmul_rn(sg, upper)),
                                              sigmoidf_ref(__fmul_rn(sg, lower))));
        const float lik = lower_bound_nan(lik_raw, lik_bound);
        if (outputs) outputs[idx] = o;
        if (z_hat) z_hat[idx] = zh;
        if (likelihood) likelihood[idx] = lik;
        acc = logf(lik);
    }
    if (bits_partials) {
        const double r = block_sum((double)acc, s_red);
        if (threadIdx.x == 0) bits_partials[(size_t)blockIdx.y * gridDim.x + blockIdx.x] = r;
    }
}

// backward of one logits chain: accumulates packed-parameter gradients, returns dout/dx * gout
__device__ __forceinline__ float eb_chain_bwd(const float* P, const EbChain& ch, float x,
                                              float gout, float* gP) {
    float gh[3];
#pragma unroll
    for (int i = 0; i < 3; ++i) {
        gP[54 + i] += gout * ch.h[3][i];
        gh[i] = gout * P[54 + i];
    }
    gP[57] += gout;
#pragma unroll
    for (int k = 3; k >= 1; --k) {
        const float* M = P + 9 + 15 * (k - 1);
        float* gM = gP + 9 + 15 * (k - 1);
        float ga[3];
#pragma unroll
        for (int j = 0; j < 3; ++j) {
            const float t = ch.t[k][j];
            gM[12 + j] += gh[j] * t;                              // d/d tanh(f)
            ga[j] = gh[j] * (1.0f + M[12 + j] * (1.0f - t * t));  // through a + f*tanh(a)
            gM[9 + j] += ga[j];                                   // bias
#pragma unroll
            for (int i = 0; i < 3; ++i) gM[3 * j + i] += ga[j] * ch.h[k - 1][i];
        }
#pragma unroll
        for (int i = 0; i < 3; ++i) gh[i] = ga[0] * M[i] + ga[1] * M[3 + i] + ga[2] * M[6 + i];
    }
    float gx = 0.0f;
#pragma unroll
    for (int j = 0; j < 3; ++j) {
        const float t = ch.t[0][j];
        gP[6 + j] += gh[j] * t;
        const float ga = gh[j] * (1.0f + P[6 + j] * (1.0f - t * t));
        gP[3 + j] += ga;
        gP[j] += ga * x;
        gx += ga * P[j];
    }
    return gx;
}

__global__ void __launch_bounds__(kEbThreads)
eb_bwd_kernel(const float* __restrict__ grad_lik, const float* __restrict__ z,
              const float* __restrict__ noise, const float* __restrict__ params,
              float* __restrict__ grad_z, float* __restrict__ grad_params, float lik_bound, int B,
              int C, int S) {
    __shared__ float P[kEbP];
    __shared__ float s_g[kEbThreads / 32][kEbP];
    const int c = blockIdx.x;
    for (int i = threadIdx.x; i < kEbP; i += blockDim.x) P[i] = params[(size_t)c * kEbP + i];
    __syncthreads();
    float gP[kEbP];
#pragma unroll
    for (int i = 0; i < kEbP; ++i) gP[i] = 0.0f;
    const int e = blockIdx.y * kEbThreads + threadIdx.x;
    if (e < B * S) {
        const int b = e / S, s = e - b * S;
        const size_t idx = ((size_t)b * C + c) * S + s;
        const float zv = __ldg(z + idx);
        const float med = P[58];
        const float o = noise ? zv + __ldg(noise + idx) : rintf(zv - med) + med;
        EbChain cl, cu;
        const float xl = o - 0.5f, xu = o + 0.5f;
        const float lower = eb_logits<true>(P, xl, &cl);
        const float upper = eb_logits<true>(P, xu, &cu);
        const float sg = -signf(lower + upper);
        const float su = sigmoidf_ref(sg * upper), sl = sigmoidf_ref(sg * lower);
        const float d = su - sl;
        float g = __ldg(grad_lik + idx);
        if (!(fabsf(d) >= lik_bound || g < 0.0f)) g = 0.0f;
        const float sd = signf(d);
        const float gU = g * sd * su * (1.0f - su) * sg;
        const float gL = -g * sd * sl * (1.0f - sl) * sg;
        float gx = eb_chain_bwd(P, cu, xu, gU, gP);
        gx += eb_chain_bwd(P, cl, xl, gL, gP);
        if (noise) {
            if (grad_z) grad_z[idx] = gx;
        } else {
            if (grad_z) grad_z[idx] = 0.0f;
            gP[58] += gx;  // round mode: d o / d median = 1
        }
    }
    if (!grad_params) return;
    // block reduction of the 59 accumulators, then one atomic per value per CTA
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int i = 0; i < kEbP - 1; ++i) {
        const float r = warp_sum(gP[i]);
        if (lane == 0) s_g[wid][i] = r;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < kEbP - 1; i += blockDim.x) {
        float r = 0.0f;
        for (int w = 0; w < kEbThreads / 32; ++w) r += s_g[w][i];
        atomicAdd(grad_params + (size_t)c * kEbP + i, r);
    }
}

// ------------------------------------------------------------------ bits finalize
__global__ void __launch_bounds__(256)
bits_finalize_kernel(const double* __restrict__ partials, const int32_t* __restrict__ seg,
                     const double* __restrict__ scales, double* __restrict__ out) {
    __shared__ double s_red[8];
    pdl_prologue();
    const int i = blockIdx.x;
    const int lo = seg[i], hi = seg[i + 1];
    double acc = 0.0;
    for (int k = lo + threadIdx.x; k < hi; k += blockDim.x) acc += partials[k];
    const double r = block_sum(acc, s_red);
    if (threadIdx.x == 0) out[i] = r * scales[i];
}

}  // namespace dsvc

using namespace dsvc;

static inline long long cdiv(long long a, long long b) { return (a + b - 1) / b; }

static bool gc_vec_ok(const GcArgs& a, long long rows) {
    if (a.inner % 4) return false;
    const void* ptrs[] = {a.x, a.scales, a.means, a.noise, a.outputs, a.likelihood,
                          a.y_hat, a.symbols, a.indexes};
    for (const void* p : ptrs)
        if (p && !aligned16(p)) return false;
    if (rows > 1 && (a.x_rs % 4 || a.scales_rs % 4 || (a.means && a.means_rs % 4) ||
                     (a.noise && a.noise_rs % 4)))
        return false;
    return true;
}

extern "C" int dsvc_reduce_slots(int64_t rows, int64_t inner) {
    if (rows <= 0 || inner <= 0) return 0;
    return (int)(rows * cdiv(inner, (long long)kGcThreads * kGcVec));
}

extern "C" int dsvc_gc_fwd_f32(const float* x, const float* scales, const float* means,
                               const float* noise, float* outputs, float* likelihood,
                               float* y_hat, int32_t* symbols, int32_t* indexes,
                               const float* scale_table, int n_table, double* bits_partials,
                               float scale_bound, float lik_bound, int64_t rows, int64_t inner,
                               int64_t x_rs, int64_t scales_rs, int64_t means_rs,
                               int64_t noise_rs, void* stream) {
    DSVC_CHECK_ARG(x && scales && rows >= 0 && inner >= 0);
    if (rows == 0 || inner == 0) return 0;
    DSVC_CHECK_ARG(rows <= 65535);
    DSVC_CHECK_ARG(!indexes || (scale_table && n_table >= 1 && n_table <= kMaxTable));
    GcArgs a{x, scales, means, noise, outputs, likelihood, y_hat, symbols, indexes, scale_table,
             n_table, bits_partials, scale_bound, lik_bound, inner, x_rs, scales_rs, means_rs,
             noise_rs};
    cudaStream_t st = (cudaStream_t)stream;
    const bool vec = gc_vec_ok(a, rows);
    dim3 grid((unsigned)cdiv(inner, (long long)kGcThreads * kGcVec), (unsigned)rows);
    prefer_max_shared_carveout(gc_fwd_kernel<true>);
    prefer_max_shared_carveout(gc_fwd_kernel<false>);
    if (vec) launch_pdl(gc_fwd_kernel<true>, grid, dim3(kGcThreads), 0, st, a);
    else launch_pdl(gc_fwd_kernel<false>, grid, dim3(kGcThreads), 0, st, a);
    DSVC_RETURN_LAST();
}

extern "C" int dsvc_gc_bwd_f32(const float* grad_lik, const float* x, const float* scales,
                               const float* means, const float* noise, float* grad_x,
                               float* grad_scales, float* grad_means, float scale_bound,
                               float lik_bound, int64_t rows, int64_t inner, int64_t x_rs,
                               int64_t scales_rs, int64_t means_rs, int64_t noise_rs,
                               void* stream) {
    DSVC_CHECK_ARG(grad_lik && x && scales && rows >= 0 && inner >= 0);
    if (rows == 0 || inner == 0) return 0;
    DSVC_CHECK_ARG(rows <= 65535);
    GcBwdArgs a{grad_lik, x, scales, means, noise, grad_x, grad_scales, grad_means, scale_bound,
                lik_bound, inner, x_rs, scales_rs, means_rs, noise_rs};
    dim3 grid((unsigned)cdiv(inner, kGcThreads), (unsigned)rows);
    gc_bwd_kernel<<<grid, kGcThreads, 0, (cudaStream_t)stream>>>(a);
    DSVC_RETURN_LAST();
}

extern "C" int dsvc_eb_reduce_slots(int B, int C, int S) {
    if (B <= 0 || C <= 0 || S <= 0) return 0;
    return (int)(C * cdiv((long long)B * S, kEbThreads));
}

extern "C" int dsvc_eb_fwd_f32(const float* z, const float* noise, const float* params,
                               float* outputs, float* likelihood, float* z_hat,
                               double* bits_partials, float lik_bound, int B, int C, int S,
                               void* stream) {
    DSVC_CHECK_ARG(z && params && B >= 0 && C >= 0 && S >= 0);
    if (B == 0 || C == 0 || S == 0) return 0;
    DSVC_CHECK_ARG((long long)B * S < (1ll << 31) && cdiv((long long)B * S, kEbThreads) <= 65535);
    dim3 grid((unsigned)C, (unsigned)cdiv((long long)B * S, kEbThreads));
    prefer_max_shared_carveout(eb_fwd_kernel);
    launch_pdl(eb_fwd_kernel, grid, dim3(kEbThreads), 0, (cudaStream_t)stream, z, noise, params, outputs, likelihood,
               z_hat, bits_partials, lik_bound, B, C, S);
    DSVC_RETURN_LAST();
}

extern "C" int dsvc_eb_bwd_f32(const float* grad_lik, const float* z, const float* noise,
                               const float* params, float* grad_z, float* grad_params,
                               float lik_bound, int B, int C, int S, void* stream) {
    DSVC_CHECK_ARG(grad_lik && z && params && B >= 0 && C >= 0 && S >= 0);
    if (B == 0 || C == 0 || S == 0) return 0;
    DSVC_CHECK_ARG((long long)B * S < (1ll << 31) && cdiv((long long)B * S, kEbThreads) <= 65535);
    dim3 grid((unsigned)C, (unsigned)cdiv((long long)B * S, kEbThreads));
    eb_bwd_kernel<<<grid, kEbThreads, 0, (cudaStream_t)stream>>>(
        grad_lik, z, noise, params, grad_z, grad_params, lik_bound, B, C, S);
    DSVC_RETURN_LAST();
}

extern "C" int dsvc_bits_finalize_f64(const double* partials, const int32_t* seg_offsets,
                                      const double* scales, double* out, int nseg, void* stream) {
    DSVC_CHECK_ARG(partials && seg_offsets && scales && out && nseg >= 0);
    if (nseg == 0) return 0;
    prefer_max_shared_carveout(bits_finalize_kernel);
    launch_pdl(bits_finalize_kernel, dim3(nseg), dim3(256), 0, (cudaStream_t)stream, partials, seg_offsets, scales, out);
    DSVC_RETURN_LAST();
}

// ------------------------------------------------------------------ bottleneck parameter packing
// [C, 60] packed parameters from the module's 15 raw tensors in ONE launch (forward) and their
// gradients in one more (backward): softplus of the matrices, the biases, tanh of the factors,
// the median (quantiles[:, 0, 1]), one pad.  The eager version is ~15 tiny launches forward and
// ~25 backward per bottleneck -- most of the kernel nodes of a training step.  Same math functions
// as ATen (softplus: x > 20 ? x : log1p(exp(x)); tanhf).
namespace dsvc {
struct EbRaw {
    const float* matrix[5];   // [C, f_{i+1}, f_i]
    const float* bias[5];     // [C, f_{i+1}, 1]
    const float* factor[4];   // [C, f_{i+1}, 1]
    const float* quantiles;   // [C, 1, 3]
};
struct EbRawGrad {
    float* matrix[5];
    float* bias[5];
    float* factor[4];
    float* quantiles;
};
// packed offsets: layer i occupies [off_i, off_i + m_i + b_i + f_i)
__device__ __forceinline__ void eb_slot(int j, int& layer, int& kind, int& k, int& width) {
    // widths: matrices 3, 9, 9, 9, 3; biases 3, 3, 3, 3, 1; factors 3, 3, 3, 3
    const int msz[5] = {3, 9, 9, 9, 3}, bsz[5] = {3, 3, 3, 3, 1};
    int off = 0;
#pragma unroll
    for (int i = 0; i < 5; ++i) {
        if (j < off + msz[i]) { layer = i; kind = 0; k = j - off; width = msz[i]; return; }
        off += msz[i];
        if (j < off + bsz[i]) { layer = i; kind = 1; k = j - off; width = bsz[i]; return; }
        off += bsz[i];
        if (i < 4) {
            if (j < off + 3) { layer = i; kind = 2; k = j - off; width = 3; return; }
            off += 3;
        }
    }
    layer = 0; kind = j == 58 ? 3 : 4; k = 0; width = 1;  // 58: median, 59: pad
}

__global__ void __launch_bounds__(64) eb_pack_kernel(EbRaw r, float* __restrict__ packed, int C) {
    const int c = blockIdx.x, j = threadIdx.x;
    if (c >= C || j >= kEbP) return;
    int layer, kind, k, width;
    eb_slot(j, layer, kind, k, width);
    float v = 0.0f;
    if (kind == 0) {
        const float x = r.matrix[layer][(size_t)c * width + k];
        v = x > 20.0f ? x : log1pf(expf(x));
    } else if (kind == 1) {
        v = r.bias[layer][(size_t)c * width + k];
    } else if (kind == 2) {
        v = tanhf(r.factor[layer][(size_t)c * width + k]);
    } else if (kind == 3) {
        v = r.quantiles[(size_t)c * 3 + 1];
    }
    packed[(size_t)c * kEbP + j] = v;
}

__global__ void __launch_bounds__(64)
eb_pack_bwd_kernel(EbRaw r, const float* __restrict__ g_packed, EbRawGrad g, int C) {
    const int c = blockIdx.x, j = threadIdx.x;
    if (c >= C || j >= kEbP) return;
    int layer, kind, k, width;
    eb_slot(j, layer, kind, k, width);
    const float gp = g_packed[(size_t)c * kEbP + j];
    if (kind == 0) {
        const float x = r.matrix[layer][(size_t)c * width + k];
        // d softplus / dx = sigmoid(x) (ATen: z = exp(x); x > threshold ? g : g * z / (z + 1))
        const float z = expf(x);
        if (g.matrix[layer]) g.matrix[layer][(size_t)c * width + k] = x > 20.0f ? gp : gp * z / (z + 1.0f);
    } else if (kind == 1) {
        if (g.bias[layer]) g.bias[layer][(size_t)c * width + k] = gp;
    } else if (kind == 2) {
        const float t = tanhf(r.factor[layer][(size_t)c * width + k]);
        if (g.factor[layer]) g.factor[layer][(size_t)c * width + k] = gp * (1.0f - t * t);
    } else if (kind == 3) {
        if (g.quantiles) {
            g.quantiles[(size_t)c * 3 + 0] = 0.0f;
            g.quantiles[(size_t)c * 3 + 1] = gp;
            g.quantiles[(size_t)c * 3 + 2] = 0.0f;
        }
    }
}
}  // namespace dsvc

extern "C" int dsvc_eb_pack_f32(const float* const* raw15, float* packed, int C, void* stream) {
    DSVC_CHECK_ARG(raw15 && packed && C >= 0);
    if (C == 0) return 0;
    dsvc::EbRaw r;
    for (int i = 0; i < 5; ++i) { r.matrix[i] = raw15[i]; r.bias[i] = raw15[5 + i]; }
    for (int i = 0; i < 4; ++i) r.factor[i] = raw15[10 + i];
    r.quantiles = raw15[14];
    for (int i = 0; i < 15; ++i) DSVC_CHECK_ARG(raw15[i] != nullptr);
    dsvc::eb_pack_kernel<<<(unsigned)C, 64, 0, (cudaStream_t)stream>>>(r, packed, C);
    DSVC_RETURN_LAST();
}

extern "C" int dsvc_eb_pack_bwd_f32(const float* const* raw15, const float* grad_packed, float* const* grad_raw15,
                                    int C, void* stream) {
    DSVC_CHECK_ARG(raw15 && grad_packed && grad_raw15 && C >= 0);
    if (C == 0) return 0;
    dsvc::EbRaw r;
    dsvc::EbRawGrad g;
    for (int i = 0; i < 5; ++i) {
        r.matrix[i] = raw15[i]; r.bias[i] = raw15[5 + i];
        g.matrix[i] = grad_raw15[i]; g.bias[i] = grad_raw15[5 + i];
    }
    for (int i = 0; i < 4; ++i) { r.factor[i] = raw15[10 + i]; g.factor[i] = grad_raw15[10 + i]; }
    r.quantiles = raw15[14];
    g.quantiles = grad_raw15[14];
    for (int i = 0; i < 15; ++i) DSVC_CHECK_ARG(raw15[i] != nullptr);
    dsvc::eb_pack_bwd_kernel<<<(unsigned)C, 64, 0, (cudaStream_t)stream>>>(r, grad_packed, g, C);
    DSVC_RETURN_LAST();
}

extern "C" int dsvc_abi_version(void) { return DSVC_ABI_VERSION; }

extern "C" const char* dsvc_error_string(int err) { return cudaGetErrorString((cudaError_t)err); }

extern "C" int dsvc_device_arch(void) {
    int dev = 0, major = 0, minor = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev) != cudaSuccess) return -1;
    if (cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev) != cudaSuccess) return -1;
    return major * 10 + minor;
}
