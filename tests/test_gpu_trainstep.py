"""BASELINE configs[2]: the training frame-step (forward + backward) of the hot path through the
drop-in API against autograd of the oracle (reference torch_warp + compressai restatement)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


def oracle_step(R, inputs, models_o, cotangents):
    """The same step on the oracle: (loss, {leaf: grad}, [pyr_flow grads], eb parameter grads)."""
    from deepsvc_b200.trainstep import LEAVES
    d = dict(inputs)
    for k in LEAVES:
        d[k] = inputs[k].detach().clone().requires_grad_(True)
    d["pyr_flow"] = [f.detach().clone().requires_grad_(True) for f in inputs["pyr_flow"]]
    B, _, H, W = inputs["ref_frame"].shape
    loss = 0.0
    outs = [R.torch_warp(img, fl) for img, fl in zip(d["pyr_img"], d["pyr_flow"])]
    outs += [R.torch_warp(d["ref_frame"], d["flow"]), R.torch_warp(d["feature"], d["flow"])]
    cots = list(cotangents["spynet"]) + [cotangents["warped_frame"], cotangents["warped_feature"]]
    for name in ("mv", "res"):
        eb, gc = models_o[name]
        for p in eb.parameters():
            p.grad = None
        _, _, y_lik, z_lik = R.codec_entropy_forward(
            eb, gc, d[f"{name}_y"], d[f"{name}_z"], d[f"{name}_scales"], d[f"{name}_means"],
            training=True, noise_y=d[f"{name}_noise_y"], noise_z=d[f"{name}_noise_z"])
        loss = loss + (R.bits_from_likelihoods(y_lik) + R.bits_from_likelihoods(z_lik)) / (B * H * W)
    torch.autograd.backward(outs + [loss], cots + [None])
    pg = {name: {n: p.grad for n, p in models_o[name][0].named_parameters() if p.grad is not None}
          for name in ("mv", "res")}
    return loss.detach(), {k: d[k].grad for k in LEAVES}, [f.grad for f in d["pyr_flow"]], pg


def _setup(oracle, B, H, W, dev, oracle_dev):
    import deepsvc_b200 as dsvc
    from deepsvc_b200 import synthetic
    from deepsvc_b200.trainstep import TrainStepHotPath, make_cotangents
    cpu_in = synthetic.make_pframe_inputs(B=B, H=H, W=W, seed=16, training=True)
    cot = make_cotangents(cpu_in)
    mo, mg = {}, {}
    for name, ch in (("mv", 64), ("res", 96)):
        eb_o, gc_o = oracle.make_entropy_models(ch, seed=ch)
        eb = dsvc.EntropyBottleneck(ch)
        eb.load_state_dict(eb_o.state_dict(), strict=False)
        mo[name] = (eb_o.to(oracle_dev).train(), gc_o.to(oracle_dev).train())
        mg[name] = (eb.to(dev).train(), dsvc.GaussianConditional(None).to(dev).train())
    ts = TrainStepHotPath(synthetic.to_device(cpu_in, dev), mg, synthetic.to_device(cot, dev))
    ref = oracle_step(oracle, synthetic.to_device(cpu_in, oracle_dev), mo, synthetic.to_device(cot, oracle_dev))
    return ts, ref


def _compare(ts, ref, tol=1e-4):
    from deepsvc_b200.trainstep import LEAVES
    loss_r, g_r, pf_r, pg_r = ref
    assert abs(float(ts.loss) - float(loss_r)) <= 1e-4 * abs(float(loss_r))
    got = ts.grads()

    def close(a, b, what):
        b = b.to(a.device)
        err = (a - b).abs().max().item()
        assert err <= tol * max(b.abs().max().item(), 1e-12), f"{what}: {err} vs scale {b.abs().max().item()}"

    for k in LEAVES:
        close(got[k], g_r[k], k)
    for i, (a, b) in enumerate(zip(got["pyr_flow"], pf_r)):
        if b.abs().max().item() == 0.0:   # coarsest level: zero flow, gradient may still be non-zero
            assert a.abs().max().item() == 0.0
        else:
            close(a, b, f"pyr_flow[{i}]")
    for name in ("mv", "res"):
        eb = ts.models[name][0]
        for n, p in eb.named_parameters():
            if n in pg_r[name]:
                close(p.grad, pg_r[name][n], f"{name}.{n}")


def test_train_step_vs_cpu_oracle(oracle):
    """Small batch against the CPU oracle (the reference's CPU branch divides the flow)."""
    import deepsvc_b200 as dsvc
    dev = torch.device("cuda:0")
    dsvc.set_flow_arithmetic("cpu")
    try:
        ts, ref = _setup(oracle, 2, 64, 128, dev, torch.device("cpu"))
        ts.step()
        torch.cuda.synchronize()
        _compare(ts, ref)
    finally:
        dsvc.set_flow_arithmetic("cuda")


def test_train_step_cfg3_vs_gpu_oracle_and_graph(oracle):
    """configs[2] at full size (B=8, 256x256) against the oracle's ops moved to the GPU (stock
    grid_sample backward), eager and as a captured CUDA graph."""
    dev = torch.device("cuda:0")
    ts, ref = _setup(oracle, 8, 256, 256, dev, dev)
    ts.step()
    torch.cuda.synchronize()
    _compare(ts, ref)
    ts.capture()
    for t in ts.leaves():
        if t.grad is not None:
            t.grad.zero_()
    ts.replay()
    torch.cuda.synchronize()
    _compare(ts, ref)
