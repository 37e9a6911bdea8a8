"""End-to-end P-frame hot path on the GPU vs the CPU oracle, eager and graph-replayed,
plus the drop-in installation into reference-shaped code."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def test_smoke_entry():
    import __graft_entry__ as g
    g.smoke()


@pytest.mark.parametrize("B,H,W", [(1, 256, 448), (2, 128, 128)])
def test_pframe_vs_oracle_and_graph(oracle, B, H, W):
    import deepsvc_b200 as dsvc
    from deepsvc_b200 import synthetic, _lib
    from deepsvc_b200.hotpath import PFrameHotPath
    dev = torch.device("cuda:0")
    cpu_in = synthetic.make_pframe_inputs(B=B, H=H, W=W, seed=16)
    mc, mg = {}, {}
    for name, ch in (("mv", 64), ("res", 96)):
        eb_o, gc_o = oracle.make_entropy_models(ch, seed=ch)
        eb = dsvc.EntropyBottleneck(ch)
        eb.load_state_dict(eb_o.state_dict(), strict=False)
        mc[name] = (eb_o.eval(), gc_o.eval())
        mg[name] = (eb.to(dev).eval(), dsvc.GaussianConditional(None).to(dev).eval())
    with torch.no_grad():
        ref = oracle.pframe_hotpath(cpu_in, mc)
    hp = PFrameHotPath(synthetic.to_device(cpu_in, dev), mg, flow_mode=_lib.FLOW_TRUE_DIVIDE)
    assert hp.n_launches == 25
    hp.run()
    torch.cuda.synchronize()
    got = hp.results()

    def check(got):
        for a, b in zip(got["spynet"] + [got["warped_frame"], got["warped_feature"]],
                        ref["spynet"] + [ref["warped_frame"], ref["warped_feature"]]):
            err = (a.cpu() - b).abs().max().item()
            assert err <= 1e-5 * max(1.0, b.abs().max().item())
        for name in ("mv", "res"):
            assert torch.equal(got[f"{name}_y_hat"].cpu(), ref[f"{name}_y_hat"])
            assert torch.equal(got[f"{name}_z_hat"].cpu(), ref[f"{name}_z_hat"])
            rb = float(ref[f"bpp_{name}"])
            assert abs(got[f"bpp_{name}"] - rb) <= 1e-4 * abs(rb)

    check(got)
    first = (got["bpp_mv"], got["bpp_res"])
    # graph replay: zero the outputs, replay, same answers bit for bit (deterministic bits)
    hp.capture()
    for t in hp.out["spynet"] + [hp.out["warped_frame"], hp.out["warped_feature"]]:
        t.zero_()
    hp.bpp.zero_()
    hp.replay()
    torch.cuda.synchronize()
    got2 = hp.results()
    check(got2)
    assert (got2["bpp_mv"], got2["bpp_res"]) == first


def test_pframe_fused_frame_warp_variant(oracle):
    """PFrameHotPath(fuse_frame_warp=True): 24 launches, outputs bit-identical to the 25-launch
    drop-in sequence (eager and graph replay)."""
    import deepsvc_b200 as dsvc
    from deepsvc_b200 import synthetic
    from deepsvc_b200.hotpath import PFrameHotPath
    dev = torch.device("cuda:0")
    cpu_in = synthetic.make_pframe_inputs(B=1, H=256, W=448, seed=16)
    mg = {name: (dsvc.EntropyBottleneck(ch).to(dev).eval(), dsvc.GaussianConditional(None).to(dev).eval())
          for name, ch in (("mv", 64), ("res", 96))}
    gin = synthetic.to_device(cpu_in, dev)
    ref = PFrameHotPath(gin, mg)
    ref.run()
    hp = PFrameHotPath(gin, mg, fuse_frame_warp=True)
    assert hp.n_launches == 24 and ref.n_launches == 25
    hp.capture()
    hp.replay()
    torch.cuda.synchronize()
    r, h = ref.results(), hp.results()
    assert torch.equal(r["warped_frame"], h["warped_frame"])
    assert torch.equal(r["warped_feature"], h["warped_feature"])
    assert (r["bpp_mv"], r["bpp_res"]) == (h["bpp_mv"], h["bpp_res"])


def test_dropin_on_reference_shaped_codec(oracle):
    """A codec with the reference's call pattern (image_model.py:151-199: EB on z, 8 slice
    GC calls fed by convs, ste_round y_hat) gives the same y_hat (bit-exact) and bpp
    (1e-4) after swap_entropy_models() as with the oracle's entropy models on the GPU."""
    import math
    import deepsvc_b200 as dsvc
    import torch.nn as nn
    dev = torch.device("cuda:0")

    class MiniCodec(nn.Module):
        def __init__(self):
            super().__init__()
            N = 16
            self.N = N
            self.g_a = nn.Conv2d(3, N, 5, 4, 2)
            self.h_a = nn.Conv2d(N, N, 3, 2, 1)
            self.h_s = nn.ConvTranspose2d(N, 2 * N, 3, 2, 1, output_padding=1)
            self.cc = nn.ModuleList(nn.Conv2d(2 * N + 2 * min(i, 4), 4, 3, 1, 1) for i in range(8))
            self.entropy_bottleneck = oracle.EntropyBottleneck(N)
            self.gaussian_conditional = oracle.GaussianConditional(None)

        def forward(self, x, ste_round):
            y = self.g_a(x)
            z = self.h_a(y)
            _, z_lik = self.entropy_bottleneck(z)
            off = self.entropy_bottleneck._get_medians()
            z_hat = ste_round(z - off) + off
            hyper = self.h_s(z_hat)
            y_hat_slices, liks = [], []
            for i, y_slice in enumerate(y.chunk(8, 1)):
                sup = torch.cat([hyper] + y_hat_slices[:4], 1)
                p = self.cc[i](sup)
                mu, scale = p[:, :2], p[:, 2:].abs() + 0.02
                _, lik = self.gaussian_conditional(y_slice, scale.contiguous(), mu.contiguous())
                liks.append(lik)
                y_hat_slices.append(ste_round(y_slice - mu) + mu)
            return torch.cat(y_hat_slices, 1), torch.cat(liks, 1), z_lik

    torch.manual_seed(0)
    m = MiniCodec().to(dev).eval()
    x = torch.rand(2, 3, 64, 96, device=dev)
    with torch.no_grad():
        y_ref, l_ref, zl_ref = m(x, oracle.ste_round)
    assert dsvc.swap_entropy_models(m) == 2
    assert isinstance(m.gaussian_conditional, dsvc.GaussianConditional)
    with torch.no_grad():
        y_got, l_got, zl_got = m(x, dsvc.ste_round)
    assert torch.equal(y_got, y_ref)
    b_ref = (torch.log(l_ref).sum() + torch.log(zl_ref).sum()).item() / -math.log(2)
    b_got = (torch.log(l_got).sum() + torch.log(zl_got).sum()).item() / -math.log(2)
    assert abs(b_got - b_ref) <= 1e-4 * abs(b_ref)
    # training mode: same noise stream as the reference (same torch generator calls)
    m.train()
    torch.manual_seed(5)
    _, l_got, zl_got = m(x, dsvc.ste_round)
    (torch.log(l_got).sum() + torch.log(zl_got).sum()).backward()
    assert m.g_a.weight.grad is not None and torch.isfinite(m.g_a.weight.grad).all()
    assert m.entropy_bottleneck._matrix0.grad is not None


def test_codec_path_bitstreams_match_oracle(oracle):
    """The real-coding call pattern of image_model.py:201-302 (EB compress/decompress,
    per-slice build_indexes + quantize("symbols"), one buffered rANS stream, slice-by-slice
    decode + dequantize): GPU symbols/indexes + C++ coder vs the oracle's CPU path +
    pure-Python coder.  Streams are byte-identical and decode to the same y_hat."""
    import deepsvc_b200 as dsvc
    from deepsvc_b200 import ans, synthetic
    from compressai import ans as oans
    dev = torch.device("cuda:0")
    C, h, w = 16, 10, 14
    # tables are evaluated on the module's device on both sides (CPU and CUDA erfc / sigmoid differ in
    # the last bit, and compressai builds them where the parameters live): the oracle runs on the GPU
    eb_o, gc_o = oracle.make_entropy_models(C, seed=11)
    eb_o, gc_o = eb_o.to(dev), gc_o.to(dev)
    eb_o.update(force=True)
    gc_o.update_scale_table(oracle.get_scale_table(), force=True)
    eb = dsvc.EntropyBottleneck(C)
    eb.load_state_dict({k: v.cpu() for k, v in eb_o.state_dict().items()
                        if k not in ("_offset", "_quantized_cdf", "_cdf_length")}, strict=False)
    eb = eb.to(dev).eval()
    eb.update(force=True)
    gc = dsvc.GaussianConditional(None).to(dev).eval()
    gc.update_scale_table(oracle.get_scale_table())
    assert torch.equal(gc.quantized_cdf.cpu(), gc_o.quantized_cdf.cpu())
    assert torch.equal(eb.quantized_cdf.cpu(), eb_o.quantized_cdf.cpu())

    g = torch.Generator().manual_seed(3)
    z = torch.randn(1, C, 3, 4, generator=g) * 4
    z[0, 0, 0, 0] = 200.0   # outside the table -> bypass coding
    z_str_o = eb_o.compress(z.to(dev))
    z_str = eb.compress(z.to(dev))
    assert z_str == z_str_o
    z_hat = eb.decompress(z_str, z.shape[-2:])
    assert torch.equal(z_hat.cpu(), eb_o.decompress(z_str_o, z.shape[-2:]).cpu())

    y, s, m = synthetic.make_latents(1, C, h, w, g)
    y[0, 1, 2, 3] = m[0, 1, 2, 3] + 3000.0   # bypass
    yd, sd, md = y.to(dev), s.to(dev), m.to(dev)
    enc_o = oans.BufferedRansEncoder()
    enc = ans.BufferedRansEncoder()
    tables = gc._cdf_tables()
    for ys, ss, ms, yo, so, mo in zip(yd.chunk(4, 1), sd.chunk(4, 1), md.chunk(4, 1),
                                      y.chunk(4, 1), s.chunk(4, 1), m.chunk(4, 1)):
        idx_o = gc_o.build_indexes(so.to(dev)).cpu()
        sym_o = gc_o.quantize(yo.to(dev), "symbols", mo.to(dev)).cpu()
        enc_o.encode_with_indexes(sym_o.reshape(-1).tolist(), idx_o.reshape(-1).tolist(),
                                  gc_o.quantized_cdf.tolist(), gc_o.cdf_length.tolist(), gc_o.offset.tolist())
        sym, idx, y_hat = gc.quantize_and_index(ys, ss, ms)
        assert torch.equal(sym.cpu(), sym_o) and torch.equal(idx.cpu(), idx_o)
        enc.encode_with_indexes(sym, idx, tables)   # device tensors -> one int32 copy each
    stream_o, stream = enc_o.flush(), enc.flush()
    assert stream == stream_o
    dec = ans.RansDecoder()
    dec.set_stream(stream)
    for ss, ms, yo, mo in zip(sd.chunk(4, 1), md.chunk(4, 1), y.chunk(4, 1), m.chunk(4, 1)):
        idx = gc.build_indexes(ss)
        rv = torch.from_numpy(dec.decode_stream_array(idx, tables)).reshape(ss.shape).to(dev)
        y_hat = gc.dequantize(rv, ms)
        assert torch.equal(y_hat.cpu(), oracle.ste_round(yo - mo) + mo)


def test_torch_ops_layer_matches_ctypes_path(oracle):
    """The TORCH_LIBRARY operator layer (csrc_torch/ops.cpp) and the ctypes binding launch the same
    kernels: identical outputs and gradients, and the ops are visible to the dispatcher."""
    import deepsvc_b200 as d
    from deepsvc_b200 import _lib, synthetic
    from deepsvc_b200.entropy import _EntropyBottleneckFn, _GaussianConditionalFn
    from deepsvc_b200.warp import _WarpFn
    ops = _lib.torch_ops()
    assert ops is not None, "libdeepsvc_b200_torch.so not built (build() compiles it)"
    assert "deepsvc_b200::torch_warp" in str(torch.ops.deepsvc_b200.torch_warp.default._schema)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(11)
    inp = torch.randn(2, 16, 64, 96, generator=g).to(dev)
    flow = synthetic.smooth_flow(2, 64, 96, g).to(dev)
    cot = torch.randn(2, 16, 64, 96, generator=g).to(dev)
    res = []
    for fn in (lambda a, b: ops.torch_warp(a, b, _lib.FLOW_MUL_RECIPROCAL), _WarpFn.apply):
        a, b = inp.clone().requires_grad_(True), flow.clone().requires_grad_(True)
        out = fn(a, b)
        out.backward(cot)
        res.append((out.detach(), a.grad, b.grad))
    assert torch.equal(res[0][0], res[1][0])
    for x, y in zip(res[0][1:], res[1][1:]):
        assert torch.allclose(x, y, rtol=1e-5, atol=1e-6)      # atomics order only
    # GaussianConditional, noise mode, all three inputs differentiable, batch-strided slices
    y, s, m = synthetic.make_latents(3, 16, 6, 10, g)
    nz = (torch.rand(3, 16, 6, 10, generator=g) - 0.5).to(dev)
    res = []
    for fn in (ops.gaussian_conditional, _GaussianConditionalFn.apply):
        yy, ss, mm = (t.to(dev).clone().requires_grad_(True) for t in (y, s, m))
        ys, sl, ms, ns = (t.chunk(2, 1)[1] for t in (yy, ss, mm, nz))
        out, lik, yhat, bits = fn(ys, sl, ms, ns, 0.11, 1e-9, True)
        (torch.log(lik).sum() + (out * 0.3).sum() + yhat.sum()).backward()
        res.append((out.detach(), lik.detach(), yhat.detach(), bits, yy.grad, ss.grad, mm.grad))
    for x, y_ in zip(res[0], res[1]):
        assert torch.equal(x, y_)
    # EntropyBottleneck, round mode, gradient to the packed parameters and z
    eb = d.EntropyBottleneck(8).to(dev)
    z = (torch.randn(2, 8, 5, 7, generator=g) * 3).to(dev)
    res = []
    for fn in (ops.entropy_bottleneck, _EntropyBottleneckFn.apply):
        zz = z.clone().requires_grad_(True)
        packed = eb.packed_params(True)
        out, lik, zhat, bits = fn(zz, packed, None, 1e-9, True)
        (torch.log(lik).sum() + zhat.sum() + out.sum()).backward()
        grads = [p.grad.clone() for p in eb.parameters()]
        for p in eb.parameters():
            p.grad = None
        res.append([out.detach(), lik.detach(), zhat.detach(), bits, zz.grad] + grads)
    for x, y_ in zip(res[0], res[1]):
        assert torch.allclose(x, y_, rtol=1e-5, atol=1e-7)


def test_frame_symbol_pipeline_streams_and_roundtrip(oracle):
    """8f-1: FrameSymbolPipeline (18 launches into one device buffer, one pinned copy, four streams
    coded by dsvc_rans_encode_many) produces exactly the bytes of the per-slice drop-in path
    (quantize_and_index + BufferedRansEncoder, EntropyBottleneck.compress) and decodes back to the
    encoder's y_hat / z_hat bit for bit."""
    import deepsvc_b200 as dsvc
    from deepsvc_b200 import ans, synthetic
    from deepsvc_b200.codec import FrameSymbolDecoder, FrameSymbolPipeline
    dev = torch.device("cuda:0")
    cpu_in = synthetic.make_pframe_inputs(B=1, H=128, W=192, seed=5)
    cpu_in["mv_y"][0, 3, 2, 1] = cpu_in["mv_means"][0, 3, 2, 1] + 4000.0      # bypass-coded symbols
    cpu_in["res_z"][0, 5, 0, 1] = -300.0
    d = synthetic.to_device(cpu_in, dev)
    models = {}
    for name, ch in (("mv", 64), ("res", 96)):
        eb_o, _ = oracle.make_entropy_models(ch, seed=ch)
        eb = dsvc.EntropyBottleneck(ch)
        eb.load_state_dict(eb_o.state_dict(), strict=False)
        eb = eb.to(dev).eval()
        eb.update(force=True)
        gc = dsvc.GaussianConditional(None).to(dev).eval()
        gc.update_scale_table(synthetic.get_scale_table())
        models[name] = (eb, gc)
    pipe = FrameSymbolPipeline(d, models, depth=2)
    for slot in (0, 1, 0):                         # slot reuse waits for the previous copy
        pipe.launch(slot, d)
    jobs = pipe.jobs(0)
    streams = ans.encode_many(jobs, threads=3)
    assert streams == ans.encode_many(pipe.jobs(1), threads=1)
    k = 0
    for name in ("mv", "res"):
        eb, gc = models[name]
        enc = ans.BufferedRansEncoder()
        for y_s, s_s, m_s in zip(d[f"{name}_y"].chunk(8, 1), d[f"{name}_scales"].chunk(8, 1), d[f"{name}_means"].chunk(8, 1)):
            sym, idx, _ = gc.quantize_and_index(y_s, s_s, m_s)
            enc.encode_with_indexes(sym, idx, gc._cdf_tables())
        assert streams[k] == enc.flush(), f"{name} y stream"
        assert [streams[k + 1]] == eb.compress(d[f"{name}_z"]), f"{name} z stream"
        k += 2
    dec = FrameSymbolDecoder(pipe, d)
    got = dec.decode(streams, [(j[1], j[2]) for j in jobs], threads=2)
    y_ref, z_ref = pipe.reconstruction(0)
    for name in ("mv", "res"):
        assert torch.equal(got[name][0], y_ref[name]) and torch.equal(got[name][1], z_ref[name])
        assert torch.equal(y_ref[name], torch.round(d[f"{name}_y"] - d[f"{name}_means"]) + d[f"{name}_means"])
        assert torch.equal(got[name][1], models[name][0].decompress([streams[2 * ("mv", "res").index(name) + 1]],
                                                                    d[f"{name}_z"].shape[-2:]))


@pytest.mark.parametrize("carry", [False, True])
def test_host_session_delivers_the_frame_to_host_memory(oracle, carry):
    """HostSession.process (the host-buffer entry point timed as `e2e`): pinned host inputs -> H2D ->
    graph -> D2H; what arrives in host memory equals the device-resident PFrameHotPath results, over
    several pipelined frames and both slots; with carry_on_device the reference's device-resident
    state (ref_frame / feature, test_video.py:368-369) is uploaded once and warped_feature stays."""
    import deepsvc_b200 as d
    from deepsvc_b200 import synthetic
    from deepsvc_b200.hotpath import HostSession, PFrameHotPath
    dev = torch.device("cuda:0")
    cpu_in = synthetic.make_pframe_inputs(B=1, H=128, W=192, seed=9)
    models = {}
    for name, ch in (("mv", 64), ("res", 96)):
        eb_o, _ = oracle.make_entropy_models(ch, seed=ch)
        eb = d.EntropyBottleneck(ch)
        eb.load_state_dict(eb_o.state_dict(), strict=False)
        models[name] = (eb.to(dev).eval(), d.GaussianConditional(None).to(dev).eval())
    hp = PFrameHotPath(synthetic.to_device(cpu_in, dev), models)
    hp.run()
    torch.cuda.synchronize()
    want = HostSession._flat_outputs(hp, ("warped_feature",) if carry else ())
    sess = HostSession(cpu_in, models, dev, carry_on_device=carry)
    full_in = sum(t.numel() * 4 for k, v in cpu_in.items() for t in (v if isinstance(v, list) else [v]))
    carried = (cpu_in["ref_frame"].numel() + cpu_in["feature"].numel()) * 4
    assert sess.h2d_bytes == full_in - (carried if carry else 0)
    slots = [sess.process() for _ in range(5)]
    assert slots == [0, 1, 0, 1, 0]
    sess.drain()
    for slot in (0, 1):
        got = sess.wait(slot)
        assert len(got) == len(want)
        for a, b in zip(got, want):
            assert torch.equal(a, b.cpu())
    b0, b1 = sess.host_bpp(0)
    assert abs(b0 + b1 - hp.results()["bpp"]) < 1e-12
