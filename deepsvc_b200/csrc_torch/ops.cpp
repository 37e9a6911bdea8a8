// torch operator layer over the C ABI of include/deepsvc_b200.h:
// TORCH_LIBRARY(deepsvc_b200, ...) ops with C++ autograd, so that an eager drop-in call costs
// one dispatcher hop instead of a Python autograd.Function + ctypes marshalling + 3-4 Python
// tensor allocations (r01: 57-69 us of host time per call for kernels that run 5 us).
//
// What each op replaces in the reference:
//   torch_warp            /root/reference/modules.py:25-62 (differentiable in both arguments)
//   gaussian_conditional  compressai GaussianConditional.forward + ste_round, image_model.py:181-183
//   entropy_bottleneck    compressai EntropyBottleneck.forward + ste_round, image_model.py:155-162
//   gc_fwd / eb_fwd       raw fused launches (quantize / build_indexes / likelihood_bits paths)
// The arithmetic lives in libdeepsvc_b200.so (hand-written sm_100a kernels); this file only
// allocates outputs on the caller's device / current stream and wires autograd.  There is no
// CPU implementation: CPU tensors raise.
#include <ATen/ATen.h>
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAGraphsC10Utils.h>
#include <c10/cuda/CUDAGuard.h>
#include <torch/autograd.h>
#include <torch/library.h>

#include <map>
#include <mutex>
#include <tuple>

#include "../../include/deepsvc_b200.h"

namespace {

using at::Tensor;
using torch::autograd::AutogradContext;
using torch::autograd::variable_list;

void check_err(int err, const char* what) {
    TORCH_CHECK(err == 0, what, " failed: CUDA error ", err, " (", dsvc_error_string(err), ")");
}

void* stream_of(const Tensor& t) { return (void*)at::cuda::getCurrentCUDAStream(t.get_device()).stream(); }

void require_cuda_f32(const char* name, const Tensor& t) {
    TORCH_CHECK(t.is_cuda(), "deepsvc_b200.", name, ": CUDA tensors required (no CPU fallback)");
    TORCH_CHECK(t.scalar_type() == at::kFloat, "deepsvc_b200.", name, ": fp32 tensors required, got ", t.scalar_type());
}

// ------------------------------------------------------------------------------- warp
struct LinKey {
    int dev, H, W;
    bool operator<(const LinKey& o) const { return std::tie(dev, H, W) < std::tie(o.dev, o.H, o.W); }
};
std::mutex g_mu;
std::map<LinKey, std::pair<Tensor, Tensor>> g_lin;          // base-grid tables (modules.py:47-50)
std::map<std::pair<int, void*>, Tensor> g_ws, g_bwd_ws;     // per (device, stream) workspaces

std::pair<Tensor, Tensor> base_grids(const Tensor& like, int H, int W) {
    std::lock_guard<std::mutex> lk(g_mu);
    const LinKey key{(int)like.get_device(), H, W};
    auto it = g_lin.find(key);
    if (it == g_lin.end()) {
        // the reference computes the base grid with CPU linspace, then copies it (modules.py:47-52)
        auto opt = at::TensorOptions().dtype(at::kFloat);
        Tensor lx = at::linspace(-1.0, 1.0, W, opt).to(like.device());
        Tensor ly = at::linspace(-1.0, 1.0, H, opt).to(like.device());
        it = g_lin.emplace(key, std::make_pair(lx, ly)).first;
    }
    return it->second;
}

struct Scales { float sx, sy, inv_sx, inv_sy; };
Scales scales_of(int H, int W) {
    Scales s;
    s.sx = (float)((W - 1.0) / 2.0);
    s.sy = (float)((H - 1.0) / 2.0);
    s.inv_sx = 1.0f / s.sx;  // ATen div_true_kernel_cuda: a * (1 / b), opmath fp32
    s.inv_sy = 1.0f / s.sy;
    return s;
}

Tensor workspace(std::map<std::pair<int, void*>, Tensor>& cache, const Tensor& like, size_t n, bool zero) {
    void* st = stream_of(like);
    const bool capturing = c10::cuda::currentStreamCaptureStatusMayInitCtx() != c10::cuda::CaptureStatus::None;
    auto opt = like.options().dtype(at::kByte);
    if (capturing) return zero ? at::zeros({(int64_t)n}, opt) : at::empty({(int64_t)n}, opt);
    std::lock_guard<std::mutex> lk(g_mu);
    auto key = std::make_pair((int)like.get_device(), st);
    auto it = cache.find(key);
    if (it == cache.end() || (size_t)it->second.numel() < n) {
        const int64_t m = (int64_t)std::max<size_t>(n, 1 << 14);
        Tensor t = zero ? at::zeros({m}, opt) : at::empty({m}, opt);
        cache[key] = t;
        return t;
    }
    return it->second;
}

void check_warp_args(const Tensor& inp, const Tensor& flow) {
    TORCH_CHECK(inp.is_cuda() && flow.is_cuda(), "deepsvc_b200.torch_warp: CUDA tensors required (no CPU fallback)");
    TORCH_CHECK(inp.device() == flow.device(), "deepsvc_b200.torch_warp: input and flow are on different devices");
    TORCH_CHECK(inp.scalar_type() == at::kFloat && flow.scalar_type() == at::kFloat, "deepsvc_b200.torch_warp: fp32 tensors required");
    TORCH_CHECK(inp.dim() == 4 && flow.dim() == 4 && flow.size(1) == 2,
                "deepsvc_b200.torch_warp: expected input [B,C,H,W] and flow [B,2,H,W]");
    TORCH_CHECK(inp.size(0) == flow.size(0) && inp.size(2) == flow.size(2) && inp.size(3) == flow.size(3),
                "deepsvc_b200.torch_warp: shape mismatch ", inp.sizes(), " vs ", flow.sizes());
}

Tensor warp_fwd(const Tensor& input, const Tensor& flow_in, int64_t flow_mode, int64_t algo) {
    check_warp_args(input, flow_in);
    Tensor flow = flow_in.contiguous();
    Tensor inp = input;
    int layout = DSVC_LAYOUT_NCHW;
    if (!inp.is_contiguous()) {
        if (inp.is_contiguous(at::MemoryFormat::ChannelsLast) && inp.size(1) % 4 == 0) layout = DSVC_LAYOUT_NHWC;
        else inp = inp.contiguous();
    }
    Tensor out = at::empty_like(inp);
    if (out.numel() == 0) return out;
    const int B = inp.size(0), C = inp.size(1), H = inp.size(2), W = inp.size(3);
    auto lin = base_grids(inp, H, W);
    const Scales s = scales_of(H, W);
    c10::cuda::CUDAGuard guard(inp.device());
    Tensor ws;
    void* wsp = nullptr;
    size_t wsn = 0;
    if (layout == DSVC_LAYOUT_NCHW) {
        ws = workspace(g_ws, inp, std::max<size_t>(dsvc_warp_workspace_bytes(B, H, W), 64), true);
        wsp = ws.data_ptr();
        wsn = ws.numel();
    }
    const int err = dsvc_warp_fwd_f32(inp.data_ptr<float>(), flow.data_ptr<float>(), out.data_ptr<float>(), B, C, H, W,
                                      lin.first.data_ptr<float>(), lin.second.data_ptr<float>(), s.sx, s.sy, s.inv_sx,
                                      s.inv_sy, (int)flow_mode, layout, (int)algo, wsp, wsn, stream_of(inp));
    if (err) {  // a failed launch may leave the scheduler words non-zero: forget the buffer
        std::lock_guard<std::mutex> lk(g_mu);
        g_ws.clear();
    }
    check_err(err, "dsvc_warp_fwd_f32");
    return out;
}

std::tuple<Tensor, Tensor> warp_bwd(const Tensor& grad_out_in, const Tensor& input, const Tensor& flow_in,
                                    bool need_input, bool need_flow, int64_t flow_mode) {
    check_warp_args(input, flow_in);
    Tensor flow = flow_in.contiguous();
    const bool nhwc = !input.is_contiguous() && input.is_contiguous(at::MemoryFormat::ChannelsLast);
    Tensor inp = input.contiguous();   // the backward kernels are NCHW
    Tensor grad_out = grad_out_in.contiguous();
    Tensor gin, gflow;
    if (need_input) gin = at::empty_like(inp);
    if (need_flow) gflow = at::empty_like(flow);
    if (inp.numel() == 0 || !(need_input || need_flow)) return {gin, gflow};
    const int B = inp.size(0), C = inp.size(1), H = inp.size(2), W = inp.size(3);
    auto lin = base_grids(inp, H, W);
    const Scales s = scales_of(H, W);
    c10::cuda::CUDAGuard guard(inp.device());
    Tensor ws;
    if (need_input) {
        size_t n = std::max<size_t>(dsvc_warp_bwd_workspace_bytes(B, H, W), 64);
        if (C >= 8) n = std::max(n, dsvc_warp_bwd_cell_workspace_bytes(B, H, W));  // the cell-order kernel's tables
        ws = workspace(g_bwd_ws, inp, n, false);
    }
    check_err(dsvc_warp_bwd_ws_f32(grad_out.data_ptr<float>(), inp.data_ptr<float>(), flow.data_ptr<float>(),
                                   need_input ? gin.data_ptr<float>() : nullptr,
                                   need_flow ? gflow.data_ptr<float>() : nullptr, B, C, H, W,
                                   lin.first.data_ptr<float>(), lin.second.data_ptr<float>(), s.sx, s.sy, s.inv_sx,
                                   s.inv_sy, (int)flow_mode, DSVC_LAYOUT_NCHW, need_input ? ws.data_ptr() : nullptr,
                                   need_input ? (size_t)ws.numel() : 0, stream_of(inp)),
              "dsvc_warp_bwd_ws_f32");
    if (need_input && nhwc) gin = gin.contiguous(at::MemoryFormat::ChannelsLast);
    return {gin, gflow};
}

struct WarpFn : public torch::autograd::Function<WarpFn> {
    static Tensor forward(AutogradContext* ctx, const Tensor& input, const Tensor& flow, int64_t flow_mode) {
        at::AutoDispatchBelowADInplaceOrView g;
        ctx->save_for_backward({input, flow});
        ctx->saved_data["flow_mode"] = flow_mode;
        return warp_fwd(input, flow, flow_mode, DSVC_WARP_AUTO);
    }
    static variable_list backward(AutogradContext* ctx, variable_list grads) {
        auto saved = ctx->get_saved_variables();
        auto r = warp_bwd(grads[0], saved[0], saved[1], ctx->needs_input_grad(0), ctx->needs_input_grad(1),
                          ctx->saved_data["flow_mode"].toInt());
        return {std::get<0>(r), std::get<1>(r), Tensor()};
    }
};

Tensor torch_warp(const Tensor& input, const Tensor& flow, int64_t flow_mode) {
    if (at::GradMode::is_enabled() && (input.requires_grad() || flow.requires_grad())) {
        check_warp_args(input, flow);
        return WarpFn::apply(input, flow, flow_mode);
    }
    return warp_fwd(input, flow, flow_mode, DSVC_WARP_AUTO);
}

// ------------------------------------------------------------------------------- GaussianConditional
// (rows, inner, row stride) of a tensor that is dense, or dense per batch row -- the memory shape
// of y.chunk(num_slices, 1) slices (image_model.py:164); anything else is copied once.
bool dense_rows(const Tensor& t) {
    if (t.is_contiguous()) return true;
    return t.dim() >= 2 && t.size(0) > 0 && t.select(0, 0).is_contiguous();
}
Tensor in_place_layout(const Tensor& t) { return (!t.defined() || dense_rows(t)) ? t : t.contiguous(); }

struct Rows { int64_t rows, inner; int64_t rs[4]; };
Rows common_rows(const Tensor* ts[4]) {
    Rows r{1, 0, {0, 0, 0, 0}};
    const Tensor& first = *ts[0];
    const int64_t n = first.numel();
    bool all_dense = true;
    for (int i = 0; i < 4; ++i) {
        if (!ts[i] || !ts[i]->defined()) continue;
        TORCH_CHECK(ts[i]->sizes() == first.sizes(), "deepsvc_b200: shape mismatch ", ts[i]->sizes(), " vs ", first.sizes());
        all_dense = all_dense && ts[i]->is_contiguous();
    }
    if (all_dense) {
        r.rows = 1;
        r.inner = n;
        for (int i = 0; i < 4; ++i) r.rs[i] = n;
        return r;
    }
    r.rows = first.size(0);
    r.inner = n / r.rows;
    for (int i = 0; i < 4; ++i) {
        if (!ts[i] || !ts[i]->defined()) continue;
        r.rs[i] = ts[i]->is_contiguous() ? r.inner : ts[i]->stride(0);
    }
    return r;
}

enum : int64_t { GC_OUTPUTS = 1, GC_LIK = 2, GC_YHAT = 4, GC_SYMBOLS = 8, GC_INDEXES = 16, GC_BITS = 32 };

// returns (outputs, likelihood, y_hat, symbols, indexes, bits_partials); entries not asked for are
// undefined (None in Python)
using Tensor6 = std::tuple<Tensor, Tensor, Tensor, Tensor, Tensor, Tensor>;
using Tensor4 = std::tuple<Tensor, Tensor, Tensor, Tensor>;
std::vector<Tensor> gc_fwd_v(const Tensor& x_in, const Tensor& scales_in, const c10::optional<Tensor>& means_in,
                           const c10::optional<Tensor>& noise_in, const c10::optional<Tensor>& scale_table_in,
                           double scale_bound, double lik_bound, int64_t want) {
    Tensor x = in_place_layout(x_in), scales = in_place_layout(scales_in);
    Tensor means = means_in.has_value() ? in_place_layout(*means_in) : Tensor();
    Tensor noise = noise_in.has_value() ? in_place_layout(*noise_in) : Tensor();
    require_cuda_f32("gaussian_conditional", x);
    require_cuda_f32("gaussian_conditional", scales);
    if (means.defined()) require_cuda_f32("gaussian_conditional", means);
    if (noise.defined()) require_cuda_f32("gaussian_conditional", noise);
    const Tensor* ts[4] = {&x, &scales, &means, &noise};
    const Rows r = common_rows(ts);
    auto fopt = x.options();
    std::vector<Tensor> out(6);
    if (want & GC_OUTPUTS) out[0] = at::empty(x.sizes(), fopt);
    if (want & GC_LIK) out[1] = at::empty(x.sizes(), fopt);
    if (want & GC_YHAT) out[2] = at::empty(x.sizes(), fopt);
    if (want & GC_SYMBOLS) out[3] = at::empty(x.sizes(), fopt.dtype(at::kInt));
    if (want & GC_INDEXES) out[4] = at::empty(x.sizes(), fopt.dtype(at::kInt));
    if (want & GC_BITS) out[5] = at::empty({(int64_t)dsvc_reduce_slots(r.rows, r.inner)}, fopt.dtype(at::kDouble));
    Tensor table;
    if (want & GC_INDEXES) {
        TORCH_CHECK_VALUE(scale_table_in.has_value() && scale_table_in->numel() >= 1,
                          "build_indexes needs a non-empty scale_table (call update_scale_table)");
        table = scale_table_in->contiguous();
        TORCH_CHECK(table.device() == x.device() && table.scalar_type() == at::kFloat,
                    "deepsvc_b200: scale_table must be fp32 on the input's device");
    }
    if (x.numel() == 0) return out;
    c10::cuda::CUDAGuard guard(x.device());
    auto fp = [](const Tensor& t) { return t.defined() ? t.data_ptr<float>() : nullptr; };
    check_err(dsvc_gc_fwd_f32(x.data_ptr<float>(), scales.data_ptr<float>(), fp(means), fp(noise), fp(out[0]), fp(out[1]),
                              fp(out[2]), out[3].defined() ? out[3].data_ptr<int32_t>() : nullptr,
                              out[4].defined() ? out[4].data_ptr<int32_t>() : nullptr, fp(table),
                              table.defined() ? (int)table.numel() : 0,
                              out[5].defined() ? out[5].data_ptr<double>() : nullptr, (float)scale_bound,
                              (float)lik_bound, r.rows, r.inner, r.rs[0], r.rs[1], means.defined() ? r.rs[2] : 0,
                              noise.defined() ? r.rs[3] : 0, stream_of(x)),
              "dsvc_gc_fwd_f32");
    return out;
}

Tensor6 gc_fwd(const Tensor& x, const Tensor& scales, const c10::optional<Tensor>& means,
               const c10::optional<Tensor>& noise, const c10::optional<Tensor>& scale_table, double scale_bound,
               double lik_bound, int64_t want) {
    auto r = gc_fwd_v(x, scales, means, noise, scale_table, scale_bound, lik_bound, want);
    return Tensor6(r[0], r[1], r[2], r[3], r[4], r[5]);
}

std::tuple<Tensor, Tensor, Tensor> gc_bwd(const Tensor& grad_lik_in, const Tensor& x, const Tensor& scales,
                                          const Tensor& means, const Tensor& noise, double scale_bound,
                                          double lik_bound, bool need_x, bool need_s, bool need_m) {
    Tensor gx, gs, gm;
    Tensor grad_lik = grad_lik_in.contiguous();
    const Tensor* ts[4] = {&x, &scales, &means, &noise};
    const Rows r = common_rows(ts);
    if (need_x) gx = at::empty(x.sizes(), x.options());
    if (need_s) gs = at::empty(x.sizes(), x.options());
    if (need_m && means.defined()) gm = at::empty(x.sizes(), x.options());
    if (x.numel() == 0) return {gx, gs, gm};
    c10::cuda::CUDAGuard guard(x.device());
    auto fp = [](const Tensor& t) { return t.defined() ? t.data_ptr<float>() : nullptr; };
    check_err(dsvc_gc_bwd_f32(grad_lik.data_ptr<float>(), x.data_ptr<float>(), scales.data_ptr<float>(), fp(means),
                              fp(noise), fp(gx), fp(gs), fp(gm), (float)scale_bound, (float)lik_bound, r.rows, r.inner,
                              r.rs[0], r.rs[1], means.defined() ? r.rs[2] : 0, noise.defined() ? r.rs[3] : 0,
                              stream_of(x)),
              "dsvc_gc_bwd_f32");
    return {gx, gs, gm};
}

Tensor acc(const Tensor& a, const Tensor& b) {
    if (!a.defined()) return b;
    if (!b.defined()) return a;
    return a + b;
}

// (outputs, likelihood, y_hat, bits_partials) with the reference's gradients: likelihood -> x,
// scales, means (LowerBound rules inside the kernel); outputs -> x (noise mode) or means (round
// mode); y_hat -> x (straight-through, image_model.py:183)
struct GaussianConditionalFn : public torch::autograd::Function<GaussianConditionalFn> {
    static variable_list forward(AutogradContext* ctx, const Tensor& x_in, const Tensor& scales_in,
                                 const c10::optional<Tensor>& means_in, const c10::optional<Tensor>& noise_in,
                                 double scale_bound, double lik_bound, bool want_bits) {
        at::AutoDispatchBelowADInplaceOrView g;
        Tensor x = in_place_layout(x_in), scales = in_place_layout(scales_in);
        Tensor means = means_in.has_value() ? in_place_layout(*means_in) : Tensor();
        Tensor noise = noise_in.has_value() ? in_place_layout(*noise_in) : Tensor();
        auto r = gc_fwd_v(x, scales, means.defined() ? c10::optional<Tensor>(means) : c10::nullopt,
                        noise.defined() ? c10::optional<Tensor>(noise) : c10::nullopt, c10::nullopt, scale_bound,
                        lik_bound, GC_OUTPUTS | GC_LIK | GC_YHAT | (want_bits ? GC_BITS : 0));
        ctx->save_for_backward({x, scales, means, noise});
        ctx->saved_data["sb"] = scale_bound;
        ctx->saved_data["lb"] = lik_bound;
        ctx->set_materialize_grads(false);
        Tensor bits = r[5].defined() ? r[5] : at::empty({0}, x.options().dtype(at::kDouble));
        ctx->mark_non_differentiable({bits});
        return {r[0], r[1], r[2], bits};
    }
    static variable_list backward(AutogradContext* ctx, variable_list grads) {
        auto saved = ctx->get_saved_variables();
        const Tensor &x = saved[0], &scales = saved[1], &means = saved[2], &noise = saved[3];
        const bool need_x = ctx->needs_input_grad(0), need_s = ctx->needs_input_grad(1);
        const bool need_m = ctx->needs_input_grad(2) && means.defined();
        Tensor gx, gs, gm;
        if (grads[1].defined() && (need_x || need_s || need_m)) {
            auto r = gc_bwd(grads[1], x, scales, means, noise, ctx->saved_data["sb"].toDouble(),
                            ctx->saved_data["lb"].toDouble(), need_x, need_s, need_m);
            gx = std::get<0>(r); gs = std::get<1>(r); gm = std::get<2>(r);
        }
        if (grads[0].defined()) {
            if (noise.defined()) { if (need_x) gx = acc(gx, grads[0]); }
            else if (need_m) gm = acc(gm, grads[0]);
        }
        if (grads[2].defined() && need_x) gx = acc(gx, grads[2]);
        return {gx, gs, gm, Tensor(), Tensor(), Tensor(), Tensor()};
    }
};

std::tuple<Tensor, Tensor, Tensor, Tensor> gaussian_conditional(const Tensor& x, const Tensor& scales,
                                                                const c10::optional<Tensor>& means,
                                                                const c10::optional<Tensor>& noise, double scale_bound,
                                                                double lik_bound, bool want_bits) {
    const bool need_grad = at::GradMode::is_enabled() &&
                           (x.requires_grad() || scales.requires_grad() || (means.has_value() && means->requires_grad()));
    if (!need_grad) {
        auto r = gc_fwd_v(x, scales, means, noise, c10::nullopt, scale_bound, lik_bound,
                          GC_OUTPUTS | GC_LIK | GC_YHAT | (want_bits ? GC_BITS : 0));
        Tensor bits = r[5].defined() ? r[5] : at::empty({0}, x.options().dtype(at::kDouble));
        return {r[0], r[1], r[2], bits};
    }
    auto r = GaussianConditionalFn::apply(x, scales, means, noise, scale_bound, lik_bound, want_bits);
    return {r[0], r[1], r[2], r[3]};
}

// ------------------------------------------------------------------------------- EntropyBottleneck
enum : int64_t { EB_OUTPUTS = 1, EB_LIK = 2, EB_ZHAT = 4, EB_BITS = 8 };

// returns (outputs, likelihood, z_hat, bits_partials)
std::vector<Tensor> eb_fwd_v(const Tensor& z_in, const Tensor& packed, const c10::optional<Tensor>& noise_in,
                           double lik_bound, int64_t want) {
    require_cuda_f32("EntropyBottleneck", z_in);
    require_cuda_f32("EntropyBottleneck", packed);
    Tensor z = z_in.contiguous();
    Tensor noise = noise_in.has_value() ? noise_in->contiguous() : Tensor();
    if (noise.defined()) require_cuda_f32("EntropyBottleneck", noise);
    TORCH_CHECK(z.dim() >= 2, "deepsvc_b200.EntropyBottleneck: expected [B, C, ...] input");
    const int B = z.size(0), C = z.size(1);
    const int S = (int)(z.numel() / std::max<int64_t>((int64_t)B * C, 1));
    TORCH_CHECK(packed.dim() == 2 && packed.size(0) == C && packed.size(1) == DSVC_EB_PARAMS_PER_CHANNEL && packed.is_contiguous(),
                "deepsvc_b200.EntropyBottleneck: channel mismatch");
    std::vector<Tensor> out(4);
    if (want & EB_OUTPUTS) out[0] = at::empty_like(z);
    if (want & EB_LIK) out[1] = at::empty_like(z);
    if (want & EB_ZHAT) out[2] = at::empty_like(z);
    if (want & EB_BITS) out[3] = at::empty({(int64_t)dsvc_eb_reduce_slots(B, C, S)}, z.options().dtype(at::kDouble));
    if (z.numel() == 0) return out;
    c10::cuda::CUDAGuard guard(z.device());
    auto fp = [](const Tensor& t) { return t.defined() ? t.data_ptr<float>() : nullptr; };
    check_err(dsvc_eb_fwd_f32(z.data_ptr<float>(), fp(noise), packed.data_ptr<float>(), fp(out[0]), fp(out[1]), fp(out[2]),
                              out[3].defined() ? out[3].data_ptr<double>() : nullptr, (float)lik_bound, B, C, S,
                              stream_of(z)),
              "dsvc_eb_fwd_f32");
    return out;
}

Tensor4 eb_fwd(const Tensor& z, const Tensor& packed, const c10::optional<Tensor>& noise, double lik_bound, int64_t want) {
    auto r = eb_fwd_v(z, packed, noise, lik_bound, want);
    return Tensor4(r[0], r[1], r[2], r[3]);
}

struct EntropyBottleneckFn : public torch::autograd::Function<EntropyBottleneckFn> {
    static variable_list forward(AutogradContext* ctx, const Tensor& z_in, const Tensor& packed_in,
                                 const c10::optional<Tensor>& noise_in, double lik_bound, bool want_bits) {
        at::AutoDispatchBelowADInplaceOrView g;
        Tensor z = z_in.contiguous(), packed = packed_in.contiguous();
        Tensor noise = noise_in.has_value() ? noise_in->contiguous() : Tensor();
        auto r = eb_fwd_v(z, packed, noise.defined() ? c10::optional<Tensor>(noise) : c10::nullopt, lik_bound,
                          EB_OUTPUTS | EB_LIK | EB_ZHAT | (want_bits ? EB_BITS : 0));
        ctx->save_for_backward({z, packed, noise});
        ctx->saved_data["lb"] = lik_bound;
        ctx->set_materialize_grads(false);
        Tensor bits = r[3].defined() ? r[3] : at::empty({0}, z.options().dtype(at::kDouble));
        ctx->mark_non_differentiable({bits});
        return {r[0], r[1], r[2], bits};
    }
    static variable_list backward(AutogradContext* ctx, variable_list grads) {
        auto saved = ctx->get_saved_variables();
        const Tensor &z = saved[0], &packed = saved[1], &noise = saved[2];
        const bool need_z = ctx->needs_input_grad(0), need_p = ctx->needs_input_grad(1);
        Tensor gz, gp;
        const int B = z.size(0), C = z.size(1);
        const int S = (int)(z.numel() / std::max<int64_t>((int64_t)B * C, 1));
        if (grads[1].defined() && (need_z || need_p) && z.numel()) {
            Tensor g = grads[1].contiguous();
            if (need_z) gz = at::empty_like(z);
            if (need_p) gp = at::zeros_like(packed);
            c10::cuda::CUDAGuard guard(z.device());
            check_err(dsvc_eb_bwd_f32(g.data_ptr<float>(), z.data_ptr<float>(),
                                      noise.defined() ? noise.data_ptr<float>() : nullptr, packed.data_ptr<float>(),
                                      need_z ? gz.data_ptr<float>() : nullptr, need_p ? gp.data_ptr<float>() : nullptr,
                                      (float)ctx->saved_data["lb"].toDouble(), B, C, S, stream_of(z)),
                      "dsvc_eb_bwd_f32");
        }
        if (grads[0].defined()) {
            if (noise.defined()) { if (need_z) gz = acc(gz, grads[0]); }
            else if (need_p) {  // round mode: d outputs / d median = 1
                Tensor gm = at::zeros_like(packed);
                gm.select(1, 58).copy_(grads[0].transpose(0, 1).reshape({C, -1}).sum(1));
                gp = acc(gp, gm);
            }
        }
        if (grads[2].defined() && need_z) gz = acc(gz, grads[2]);  // straight-through (image_model.py:160-162)
        return {gz, gp, Tensor(), Tensor(), Tensor()};
    }
};

std::tuple<Tensor, Tensor, Tensor, Tensor> entropy_bottleneck(const Tensor& z, const Tensor& packed,
                                                              const c10::optional<Tensor>& noise, double lik_bound,
                                                              bool want_bits) {
    const bool need_grad = at::GradMode::is_enabled() && (z.requires_grad() || packed.requires_grad());
    if (!need_grad) {
        auto r = eb_fwd_v(z, packed.contiguous(), noise, lik_bound, EB_OUTPUTS | EB_LIK | EB_ZHAT | (want_bits ? EB_BITS : 0));
        Tensor bits = r[3].defined() ? r[3] : at::empty({0}, z.options().dtype(at::kDouble));
        return {r[0], r[1], r[2], bits};
    }
    auto r = EntropyBottleneckFn::apply(z, packed, noise, lik_bound, want_bits);
    return {r[0], r[1], r[2], r[3]};
}

int64_t abi_version() { return dsvc_abi_version(); }

}  // namespace

TORCH_LIBRARY(deepsvc_b200, m) {
    m.def("abi_version() -> int", &abi_version);
    m.def("warp_fwd(Tensor input, Tensor flow, int flow_mode, int algo) -> Tensor", &warp_fwd);
    m.def("warp_bwd(Tensor grad_out, Tensor input, Tensor flow, bool need_input, bool need_flow, int flow_mode) -> (Tensor, Tensor)",
          &warp_bwd);
    m.def("torch_warp(Tensor input, Tensor flow, int flow_mode) -> Tensor", &torch_warp);
    m.def("gc_fwd(Tensor x, Tensor scales, Tensor? means, Tensor? noise, Tensor? scale_table, float scale_bound, "
          "float lik_bound, int want) -> (Tensor, Tensor, Tensor, Tensor, Tensor, Tensor)", &gc_fwd);
    m.def("gaussian_conditional(Tensor x, Tensor scales, Tensor? means, Tensor? noise, float scale_bound, "
          "float lik_bound, bool want_bits) -> (Tensor, Tensor, Tensor, Tensor)", &gaussian_conditional);
    m.def("eb_fwd(Tensor z, Tensor packed, Tensor? noise, float lik_bound, int want) -> (Tensor, Tensor, Tensor, Tensor)", &eb_fwd);
    m.def("entropy_bottleneck(Tensor z, Tensor packed, Tensor? noise, float lik_bound, bool want_bits) -> "
          "(Tensor, Tensor, Tensor, Tensor)", &entropy_bottleneck);
}
