/*
 * deepsvc_b200 -- C ABI of the B200 (sm_100a) warp + entropy P-frame hot path.
 *
 * Drop-in boundary for the DeepSVC structure/texture layer hot path.  The
 * reference (LHB116/DeepSVC) is pure Python and has no FFI of its own; every
 * entry point below names the reference callable whose arithmetic it replaces
 * (file:line into the reference tree).  The Python host side
 * (deepsvc_b200/*.py) binds these symbols with ctypes and mirrors the
 * reference's signatures (torch_warp, GaussianConditional, EntropyBottleneck,
 * ste_round); see INTEGRATION.md for the binding a maintainer would add.
 *
 * Conventions
 *   - all pointers are DEVICE pointers unless the function name ends in _host;
 *   - tensors are dense fp32, NCHW ("layout 0") unless stated; inputs are
 *     borrowed and never written; outputs are caller-allocated;
 *   - every launcher enqueues on `stream` (a cudaStream_t passed as void*),
 *     performs no allocation and no synchronisation -> CUDA-graph capturable;
 *   - return value: 0 on success, otherwise a cudaError_t value
 *     (dsvc_error_string() gives the text); DSVC_ERR_INVALID_ARG (= 1 =
 *     cudaErrorInvalidValue) for bad shapes / null pointers / misalignment.
 *     There is no CPU fallback anywhere behind this ABI.
 */
#ifndef DEEPSVC_B200_H_
#define DEEPSVC_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define DSVC_ABI_VERSION 1
#define DSVC_ERR_INVALID_ARG 1

/* ---- flow scaling arithmetic (modules.py:36-37 CPU branch, :54-55 CUDA branch) ----
 * The reference divides the pixel flow by (W-1)/2 and (H-1)/2 with a python
 * scalar.  ATen executes that as a true division on CPU and as a multiply by
 * the fp32 reciprocal on CUDA; the two differ in the last bit, which moves the
 * sampling coordinate by up to 1e-4 px at 1080p.  The caller picks which of
 * the reference's two branches to reproduce. */
#define DSVC_FLOW_MUL_RECIPROCAL 0 /* reference CUDA branch (default for the drop-in) */
#define DSVC_FLOW_TRUE_DIVIDE 1    /* reference CPU branch */

/* ---- warp algorithm selector ---- */
#define DSVC_WARP_AUTO 0   /* TMA-staged tiles when shape/alignment allow, else gather */
#define DSVC_WARP_GATHER 1 /* direct read-only-path gather */
#define DSVC_WARP_TMA 2    /* force shared-memory/TMA staging (error if unsupported) */

#define DSVC_LAYOUT_NCHW 0
#define DSVC_LAYOUT_NHWC 1 /* torch.channels_last */

int dsvc_abi_version(void);
const char* dsvc_error_string(int err);
/* Compute capability of the current device as major*10+minor (100 on B200), <0 on error. */
int dsvc_device_arch(void);

/* Backward bilinear warp, border clamp, align_corners=True.
 * Replaces modules.py:25-62 torch_warp(tensorInput, tensorFlow):
 *   out[b,c,y,x] = bilinear(input[b,c], ix, iy),
 *   gx = lin_x[x] + flow[b,0,y,x] / sx      (sx = (W-1)/2, see DSVC_FLOW_*)
 *   ix = clamp(((gx + 1) / 2) * (W-1), 0, W-1)            (same for y)
 * lin_x[W], lin_y[H]: the reference's torch.linspace(-1,1,W|H) base grids
 * (modules.py:47-50), passed in so that the CPU-computed table is reproduced
 * bit for bit.  flow is always NCHW [B,2,H,W] (ch0 = dx, ch1 = dy, pixels).
 * input/out: [B,C,H,W] in `layout`.
 *
 * workspace (device, 16-byte aligned, nullable): scheduler state (a work-unit counter)
 * of at least dsvc_warp_workspace_bytes(B, H, W) bytes for the persistent
 * shared-memory/TMA-staged kernel.  It MUST be zero-filled before its first use; every
 * launch leaves it zero-filled again, so one buffer can serve any number of launches
 * that are ordered on one stream (two launches that may run concurrently need two
 * buffers).  Tiles whose source bounding box cannot be staged are re-staged as
 * quadrants or gathered directly inside the same launch.  Without a workspace
 * DSVC_WARP_AUTO uses the gather kernel and DSVC_WARP_TMA fails with
 * DSVC_ERR_INVALID_ARG.  Results are bit-identical on every path. */
int dsvc_warp_fwd_f32(const float* input, const float* flow, float* out,
                      int B, int C, int H, int W,
                      const float* lin_x, const float* lin_y,
                      float sx, float sy, float inv_sx, float inv_sy,
                      int flow_mode, int layout, int algo,
                      void* workspace, size_t workspace_bytes, void* stream);
size_t dsvc_warp_workspace_bytes(int B, int H, int W);

/* Two tensors warped by ONE flow in one launch: out_a = torch_warp(input_a, flow),
 * out_b = torch_warp(input_b, flow), bit-identical to two dsvc_warp_fwd_f32 calls.
 * In DeepSVC.forward the frame warp of video_model.py:37 and the feature warp of
 * modules.py:429 share recon_mv: the frame's 3 planes ride on the persistent staged kernel of
 * the 64-ch feature warp (same source coordinates and staging boxes, a second tensor map)
 * instead of paying a second few-channel launch.  NCHW fp32, both [B,C*,H,W]; shapes the
 * staged kernel cannot take (Ca < 8, W % 4 != 0, small images) run as two ordinary launches. */
int dsvc_warp_fwd2_f32(const float* input_a, const float* input_b, const float* flow,
                       float* out_a, float* out_b, int B, int Ca, int Cb, int H, int W,
                       const float* lin_x, const float* lin_y,
                       float sx, float sy, float inv_sx, float inv_sy, int flow_mode,
                       void* workspace, size_t workspace_bytes, void* stream);

/* Gradient of the above (autograd of modules.py:25-62 = ATen
 * grid_sampler_2d_backward + the division by sx/sy).
 * grad_input [B,C,H,W] (nullable) MUST be zero-filled by the caller: contributions are
 * added with global reductions (per tile after combining in shared memory, or per tap).
 * grad_flow [B,2,H,W] (nullable) is overwritten.  Kernel choice: dsvc_set_warp_bwd_algo. */
int dsvc_warp_bwd_f32(const float* grad_out, const float* input, const float* flow,
                      float* grad_input, float* grad_flow,
                      int B, int C, int H, int W,
                      const float* lin_x, const float* lin_y,
                      float sx, float sy, float inv_sx, float inv_sy,
                      int flow_mode, int layout, void* stream);

/* The same gradient with a caller-provided workspace: NEITHER output needs initialising.
 * Kernel choice (DSVC_WARP_BWD_AUTO):
 *  - C >= 8, grad_input wanted and workspace_bytes >= dsvc_warp_bwd_cell_workspace_bytes(B,H,W): the
 *    cell-order kernel (csrc/warp_bwd_cell.cu).  A table launch files every output pixel under the
 *    cell (floor of its clamped source coordinate) it samples; one warp per 31 x 2 block of grad_input
 *    then walks the cells: per channel, loads from a TMA-staged box of grad_out and the input rows,
 *    shuffles, plain row stores -- no atomics on the main path, no zero-fill, any flow (cells with
 *    more than two pixels and wild regions take reduction / load paths inside the same launches).
 *  - otherwise grad_input is zero-filled here and, when the shape is eligible for the staged kernel
 *    and workspace_bytes >= dsvc_warp_bwd_workspace_bytes, the kernel is chosen PER LAUNCH on the
 *    device: a scout launch samples 32 pixels of every 64 x 16 tile of the flow, counts the tiles whose
 *    source footprint fits the staged kernel's 96 x 32 box and writes one decision word into the
 *    workspace; the staged and the per-pixel launch that follow both read it and the one not chosen
 *    exits at once (no host synchronisation, graph-capturable).
 *  - without a workspace this is dsvc_warp_bwd_f32 after a zero-fill.
 * With dsvc_set_warp_bwd_algo(DSVC_WARP_BWD_GATHER), a `workspace` (>= dsvc_warp_bwd_workspace_bytes
 * (B,H,W) bytes, any contents; one flag byte per 64 x 16 tile, written before it is read) and an
 * eligible shape (grad_input wanted, W % 4 == 0, 16-byte aligned pointers) grad_input is
 * produced by the destination-owned gather kernel (csrc/warp_bwd_gather.cu); an ineligible shape
 * is cudaErrorInvalidValue.  DSVC_WARP_BWD_CELL forces the cell-order kernel (cudaErrorInvalidValue
 * when the workspace is too small).  The workspace may hold anything on entry and is scratch. */
int dsvc_warp_bwd_ws_f32(const float* grad_out, const float* input, const float* flow,
                         float* grad_input, float* grad_flow,
                         int B, int C, int H, int W,
                         const float* lin_x, const float* lin_y,
                         float sx, float sy, float inv_sx, float inv_sy,
                         int flow_mode, int layout,
                         void* workspace, size_t workspace_bytes, void* stream);
size_t dsvc_warp_bwd_workspace_bytes(int B, int H, int W);
/* Workspace that enables the cell-order backward kernel (csrc/warp_bwd_cell.cu): the per-cell pixel
 * tables, 28 bytes per pixel plus the regions' buckets (76 MB at 1 x 1088 x 1920). */
size_t dsvc_warp_bwd_cell_workspace_bytes(int B, int H, int W);

/* Fusions around the few-channel (C <= 4) warps -- SURVEY.md 8f-3, inference only.
 * Exactly one of `flow` [B,2,H,W] and `flow_coarse` [B,2,H/2,W/2] is given.
 *  - flow_coarse: the SpyNet level of modules.py:163-168.  flow_up [B,2,H,W] is written
 *    with bilinearupsacling(flow_coarse) * 2.0 (F.interpolate x2, bilinear,
 *    align_corners=False; modules.py:107-112) and `out` = torch_warp(input, flow_up).
 *  - target [B,C,H,W] (nullable): video_model.py:37-38.  sq_partials
 *    [dsvc_warp_fused_slots(B,H,W)] receives per-CTA sums of (out - target)^2 in a fixed
 *    order; dsvc_bits_finalize_f64 with scale 1/(B*C*H*W) turns them into warp_loss. */
int dsvc_warp_fused_f32(const float* input, const float* flow, const float* flow_coarse,
                        float* flow_up, const float* target, double* sq_partials, float* out,
                        int B, int C, int H, int W, const float* lin_x, const float* lin_y,
                        float sx, float sy, float inv_sx, float inv_sy, int flow_mode,
                        void* stream);
int dsvc_warp_fused_slots(int B, int H, int W);

/* out = weight * warped + (1 - weight) * pred over n contiguous floats: the motion-compensation
 * blend of modules.py:436 (`w * warped + (1 - w) * self.out_conv(up_out)`) in one pass,
 * same order of fp32 operations as the reference's expression (bit-identical). */
int dsvc_blend_f32(const float* weight, const float* warped, const float* pred, float* out,
                   int64_t n, void* stream);

/* out = y_hat + 0.5 * tanh(lrp) over n contiguous floats: the latent-residual-prediction add of
 * image_model.py:185-188 (`lrp = 0.5 * torch.tanh(lrp); y_hat_slice += lrp`) in one pass instead of
 * three, bit-identical (same fp32 operations in the same order).  out may alias y_hat (the
 * reference adds in place).  dsvc_lrp_add_bwd_f32: grad_lrp = grad_out * 0.5 * (1 - tanh(lrp)^2)
 * (the gradient with respect to y_hat is grad_out itself). */
int dsvc_lrp_add_f32(const float* y_hat, const float* lrp, float* out, int64_t n, void* stream);
int dsvc_lrp_add_bwd_f32(const float* grad_out, const float* lrp, float* grad_lrp, int64_t n, void* stream);

/* Kernel choice of dsvc_warp_bwd_f32 (process-wide; tests and profiling).
 * DSVC_WARP_BWD_AUTO (0, default): the shared-memory staged kernel
 * (per-tile transposed-warp CSR, run sums in a shared-memory out-box, row-contiguous RED.ADD.v4.F32
 * into grad_input; csrc/warp_bwd_staged.cu)
 * when C >= 8, W % 4 == 0, W >= 64, H >= 16 and the pointers are 16-byte aligned, else the
 * per-pixel RED.ADD kernel; DSVC_WARP_BWD_DIRECT (1): always the per-pixel kernel;
 * DSVC_WARP_BWD_STAGED (2): the staged kernel or cudaErrorInvalidValue. */
#define DSVC_WARP_BWD_AUTO 0
#define DSVC_WARP_BWD_DIRECT 1
#define DSVC_WARP_BWD_STAGED 2
#define DSVC_WARP_BWD_GATHER 3 /* dsvc_warp_bwd_ws_f32 only */
#define DSVC_WARP_BWD_CELL 4   /* dsvc_warp_bwd_ws_f32 only */
int dsvc_set_warp_bwd_algo(int algo);

/* Number of double partial sums a gc launch over rows x inner elements writes. */
int dsvc_reduce_slots(int64_t rows, int64_t inner);

/* Fused Gaussian-conditional quantise / likelihood / bit estimate.
 * Replaces compressai 1.2.1 GaussianConditional.forward / .quantize /
 * .build_indexes and ops.ste_round as called from image_model.py:181,183,
 * 237-238 and the log-sum of video_model.py:39-42.  For every element:
 *   s      = max(scales, scale_bound)                      (LowerBound 0.11)
 *   q      = rint(x - means)                               (half-to-even)
 *   y_hat  = q + means                   -> y_hat   (ste_round value, :183)
 *   o      = noise ? x + noise : y_hat   -> outputs (forward()'s first result)
 *   v      = |o - means|
 *   lik    = max(.5 erfc(-(.5-v)/(s sqrt2)) - .5 erfc(-(-.5-v)/(s sqrt2)), lik_bound)
 *   symbols= (int32) q ; indexes = #{k < n_table-1 : scale_table[k] < s}
 *   bits_partials[cta] = sum over the CTA's elements of ln(lik)   (double)
 * Any output pointer may be NULL.  means may be NULL (treated as 0).
 * noise != NULL selects training ("noise") mode for outputs/likelihood.
 * Inputs are `rows` rows of `inner` contiguous floats with per-tensor row strides
 * (*_rs, in elements): this is exactly the memory shape of the reference's
 * y.chunk(num_slices, 1) slices for batch > 1 (image_model.py:164).  A dense
 * tensor is rows = 1, inner = numel.  Outputs are dense [rows, inner]. */
int dsvc_gc_fwd_f32(const float* x, const float* scales, const float* means,
                    const float* noise,
                    float* outputs, float* likelihood, float* y_hat,
                    int32_t* symbols, int32_t* indexes,
                    const float* scale_table, int n_table,
                    double* bits_partials,
                    float scale_bound, float lik_bound,
                    int64_t rows, int64_t inner,
                    int64_t x_rs, int64_t scales_rs, int64_t means_rs, int64_t noise_rs,
                    void* stream);

/* Gradient of the likelihood branch of the above (autograd of
 * GaussianConditional._likelihood + the two LowerBound rules,
 * backward pass-through if (x >= bound) or (grad < 0)).
 * grad_lik: dL/d likelihood.  Outputs (nullable): grad_x (= -grad_means in
 * noise mode, zero in round mode), grad_scales, grad_means. */
int dsvc_gc_bwd_f32(const float* grad_lik, const float* x, const float* scales,
                    const float* means, const float* noise,
                    float* grad_x, float* grad_scales, float* grad_means,
                    float scale_bound, float lik_bound,
                    int64_t rows, int64_t inner,
                    int64_t x_rs, int64_t scales_rs, int64_t means_rs, int64_t noise_rs,
                    void* stream);

/* Number of packed floats per channel expected by dsvc_eb_*: softplus'd
 * matrices, biases, tanh'd factors of the 1-3-3-3-3-1 network, then the
 * median; see deepsvc_b200/entropy.py::pack_bottleneck_params. */
#define DSVC_EB_PARAMS_PER_CHANNEL 60

/* The packed parameters of dsvc_eb_*: [C, DSVC_EB_PARAMS_PER_CHANNEL] from the module's 15 raw
 * tensors in one launch, and the gradients of the raw tensors from the packed gradient in one
 * more (compressai EntropyBottleneck._logits_cumulative applies softplus to `_matrix{i}` and tanh to
 * `_factor{i}` on every call; `_get_medians()` = quantiles[:, :, 1:2]).  raw15 / grad_raw15: HOST
 * arrays of 15 DEVICE pointers in the order _matrix0..4 ([C,f_{i+1},f_i], f = 1,3,3,3,3,1), _bias0..4
 * ([C,f_{i+1},1]), _factor0..3 ([C,f_{i+1},1]), quantiles ([C,1,3]); a NULL grad_raw15 entry is
 * skipped.  Contiguous fp32. */
int dsvc_eb_pack_f32(const float* const* raw15, float* packed, int C, void* stream);
int dsvc_eb_pack_bwd_f32(const float* const* raw15, const float* grad_packed,
                         float* const* grad_raw15, int C, void* stream);

/* Fused factorised-prior quantise / likelihood / bit estimate on z [B,C,S]
 * (S = h*w, NCHW).  Replaces compressai 1.2.1 EntropyBottleneck.forward
 * (call site image_model.py:155) and the z part of image_model.py:160-162:
 *   o   = noise ? z + noise : rint(z - med_c) + med_c
 *   L,U = logits_cumulative_c(o -/+ .5) ; sg = -sign(L+U)
 *   lik = max(|sigmoid(sg U) - sigmoid(sg L)|, lik_bound)
 * outputs <- o, z_hat <- rint(z - med)+med (always), likelihood, bits partials
 * (sum of ln lik per CTA).  Any output may be NULL. */
int dsvc_eb_fwd_f32(const float* z, const float* noise, const float* params,
                    float* outputs, float* likelihood, float* z_hat,
                    double* bits_partials, float lik_bound,
                    int B, int C, int S, void* stream);
int dsvc_eb_reduce_slots(int B, int C, int S);

/* Gradient of the likelihood branch of dsvc_eb_fwd_f32: grad_z [B,C,S]
 * (zero in round mode) and grad_params [C,DSVC_EB_PARAMS_PER_CHANNEL] w.r.t. the
 * PACKED (already softplus'd / tanh'd) parameters; must be zero-filled by the
 * caller (accumulated with atomics).  Either may be NULL. */
int dsvc_eb_bwd_f32(const float* grad_lik, const float* z, const float* noise,
                    const float* params, float* grad_z, float* grad_params,
                    float lik_bound, int B, int C, int S, void* stream);

/* out[i] = scale[i] * sum(partials[seg_offsets[i] .. seg_offsets[i+1])), fixed
 * summation order, fp64.  With scale = -1/(ln2 * pixels) this is the bpp of
 * video_model.py:39-42.  seg_offsets is a DEVICE int32 array [nseg+1]. */
int dsvc_bits_finalize_f64(const double* partials, const int32_t* seg_offsets,
                           const double* scales, double* out, int nseg, void* stream);

/* ------------------------------------------------------------------------------
 * Host-side range coder ("next" rows f-1 / f-2 of SURVEY.md 8f).  HOST pointers.
 * Replaces the native extension of the reference's dependency (compressai 1.2.1
 * cpp_exts/rans/rans_interface.cpp + cpp_exts/ops/ops.cpp) as used at
 * image_model.py:217-221,253-254,266-274,288,319-324.  Same wire format (rANS64,
 * 32-bit words, 16-bit precision, 4-bit bypass), so streams interoperate.
 * cdfs is a row-major int32 table [n_cdfs, cdf_stride]; cdf_sizes / offsets [n_cdfs]. */

/* compressai._CXX.pmf_to_quantized_cdf: pmf[n] -> cdf_out[n+1], strictly increasing,
 * cdf_out[n] = 1 << precision. */
int dsvc_pmf_to_quantized_cdf_host(const float* pmf, int n, int precision, int32_t* cdf_out);

/* compressai.ans.BufferedRansEncoder: push any number of (symbols, indexes) runs, then
 * flush once (image_model.py:221,253-254). */
void* dsvc_rans_encoder_create(void);
void dsvc_rans_encoder_destroy(void* enc);
int dsvc_rans_encoder_push(void* enc, const int32_t* symbols, const int32_t* indexes, int64_t n,
                           const int32_t* cdfs, int n_cdfs, int cdf_stride,
                           const int32_t* cdf_sizes, const int32_t* offsets);
int64_t dsvc_rans_encoder_bound(void* enc);
int dsvc_rans_encoder_flush(void* enc, uint8_t* out, int64_t out_cap, int64_t* out_len);

/* compressai.ans.RansDecoder: set_stream once, decode_stream per slice
 * (image_model.py:273-274,288). */
void* dsvc_rans_decoder_create(const uint8_t* stream, int64_t len);
void dsvc_rans_decoder_destroy(void* dec);
int dsvc_rans_decoder_decode(void* dec, const int32_t* indexes, int64_t n, const int32_t* cdfs,
                             int n_cdfs, int cdf_stride, const int32_t* cdf_sizes,
                             const int32_t* offsets, int32_t* out_symbols);

/* Many independent streams coded in parallel on `n_threads` host threads (stream s is what
 * one BufferedRansEncoder / RansDecoder would produce / consume: byte-identical).  A single
 * rANS stream is sequential; the parallelism of a byte-compatible coder is across streams --
 * mv / res codecs, y / z, and the frames in flight (image_model.py:217-221,253-254 once per
 * codec and frame).  All arrays have n_streams entries; tables are per stream.  out[s] must
 * hold (counts[s] * 3 + 2) * 4 bytes in the worst case (every symbol bypass-coded); out_len[s]
 * receives the stream length. */
int dsvc_rans_encode_many(const int32_t* const* symbols, const int32_t* const* indexes,
                          const int64_t* counts, int n_streams,
                          const int32_t* const* cdfs, const int32_t* n_cdfs,
                          const int32_t* cdf_strides, const int32_t* const* cdf_sizes,
                          const int32_t* const* offsets,
                          uint8_t* const* out, const int64_t* out_cap, int64_t* out_len,
                          int n_threads);
int dsvc_rans_decode_many(const uint8_t* const* streams, const int64_t* stream_len,
                          const int32_t* const* indexes, const int64_t* counts, int n_streams,
                          const int32_t* const* cdfs, const int32_t* n_cdfs,
                          const int32_t* cdf_strides, const int32_t* const* cdf_sizes,
                          const int32_t* const* offsets,
                          int32_t* const* out, int n_threads);

#ifdef __cplusplus
}
#endif
#endif /* DEEPSVC_B200_H_ */
