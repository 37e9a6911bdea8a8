"""One training frame-step (forward + backward) of the warp + entropy hot path.

BASELINE.json configs[2]: batch 8 of 256x256 crops, ``Learner.py:1306-1343`` in training
mode -- ``GaussianConditional`` / ``EntropyBottleneck`` in noise mode with their likelihood
tensors materialised for autograd, the 64-ch feature warp differentiated in both arguments,
the 3-ch warps (frames do not require grad) in the flow only.  The conv transforms are outside
the path: their outputs are synthetic leaves, and the distortion gradient that would reach
the warped tensors is a fixed random cotangent.

The step runs through the package's public drop-in API (``torch_warp``,
``GaussianConditional.forward_fused``, ``EntropyBottleneck.forward_fused``) and torch autograd,
exactly as the reference's modules would call it, and can be captured into a CUDA graph:

    backward of (bpp_mv + bpp_res) and of every warp output against its cotangent,
    bpp = sum(log lik) / (-ln 2 * B*H*W)

(``video_model.py:39-42,53-56,69``).  ``tests/test_gpu_trainstep.py`` runs the same step on
the oracle's ops and compares every gradient.
"""
import math

import torch

from .entropy import EntropyBottleneck, GaussianConditional
from .warp import torch_warp

NUM_SLICES = 8
LEAVES = ("feature", "flow", "mv_y", "mv_scales", "mv_means", "mv_z", "res_y", "res_scales", "res_means",
          "res_z")


def make_cotangents(inputs: dict, seed: int = 7) -> dict:
    """Fixed random d(loss)/d(warp output) tensors (CPU), scaled like an MSE gradient."""
    g = torch.Generator().manual_seed(seed)
    n = inputs["ref_frame"].numel()
    cot = {"spynet": [torch.randn(t.shape, generator=g) / n for t in inputs["pyr_img"]],
           "warped_frame": torch.randn(inputs["ref_frame"].shape, generator=g) / n,
           "warped_feature": torch.randn(inputs["feature"].shape, generator=g) / n}
    return cot


def train_algorithmic_bytes(B, H, W, feature_ch=64, mv_ch=64, res_ch=96):
    """SURVEY.md section 8d, configs[2]: forward with likelihoods materialised + backward."""
    def wf(C, h, w):
        return 4 * B * h * w * (2 * C + 2)

    def wb(C, h, w, both):
        return 4 * B * h * w * ((3 * C + 4) if both else (2 * C + 4))
    ny = B * (mv_ch + res_ch) * (H // 16) * (W // 16)
    nz = B * (mv_ch + res_ch) * (H // 64) * (W // 64)
    fwd = sum(wf(3, H >> k, W >> k) for k in range(4)) + wf(3, H, W) + wf(feature_ch, H, W) + 24 * ny + 12 * nz
    bwd = sum(wb(3, H >> k, W >> k, False) for k in range(4)) + wb(3, H, W, False) + \
        wb(feature_ch, H, W, True) + 32 * ny + 16 * nz
    return {"forward": fwd, "backward": bwd, "total": fwd + bwd,
            "feature_bwd": wb(feature_ch, H, W, True), "feature_fwd": wf(feature_ch, H, W)}


class TrainStepHotPath:
    def __init__(self, inputs: dict, models: dict, cotangents: dict):
        """inputs: ``synthetic.make_pframe_inputs(training=True)`` on one CUDA device;
        models: {"mv": (EntropyBottleneck, GaussianConditional), "res": ...} in train mode;
        cotangents: ``make_cotangents`` on the same device."""
        self.inp = dict(inputs)
        self.models = models
        self.cot = cotangents
        for name in ("mv", "res"):
            eb, gc = models[name]
            assert isinstance(eb, EntropyBottleneck) and isinstance(gc, GaussianConditional)
        for k in LEAVES:
            self.inp[k] = inputs[k].detach().clone().requires_grad_(True)
        self.inp["pyr_flow"] = [f.detach().clone().requires_grad_(True) for f in inputs["pyr_flow"]]
        B, _, H, W = inputs["ref_frame"].shape
        self.pixels = B * H * W
        self.params = [p for name in ("mv", "res") for p in models[name][0].parameters() if p.requires_grad]
        self._graph = None
        self._streams = None
        # leaf gradients are accumulated on the side streams on purpose (see forward())
        if hasattr(torch.autograd.graph, "set_warn_on_accumulate_grad_stream_mismatch"):
            torch.autograd.graph.set_warn_on_accumulate_grad_stream_mismatch(False)
        self.loss = None
        self.bpp = None

    # ------------------------------------------------------------------ one step
    def leaves(self):
        return [self.inp[k] for k in LEAVES] + list(self.inp["pyr_flow"])

    def forward(self):
        """Returns (warp outputs in cotangent order, bpp loss, {name: bpp}).

        The four independent parts of the path (feature warp | 3-ch warps | mv entropy | res
        entropy) are issued on four side streams forked from the current one; autograd runs
        each backward node on its forward stream, so a captured step is a DAG, not a chain."""
        d = self.inp
        dev = d["feature"].device
        main = torch.cuda.current_stream(dev)
        if self._streams is None:
            self._streams = [torch.cuda.Stream(dev) for _ in range(4)]
        s_feat, s_w3, s_mv, s_res = self._streams
        for st in self._streams:
            st.wait_stream(main)
        scale = -1.0 / (math.log(2) * self.pixels)
        with torch.cuda.stream(s_feat):
            feat = torch_warp(d["feature"], d["flow"])                                # modules.py:429
        with torch.cuda.stream(s_w3):
            outs = [torch_warp(img, fl) for img, fl in zip(d["pyr_img"], d["pyr_flow"])]  # modules.py:167
            outs.append(torch_warp(d["ref_frame"], d["flow"]))                        # video_model.py:37
        outs.append(feat)
        bpp = {}
        for name, st in (("mv", s_mv), ("res", s_res)):
            with torch.cuda.stream(st):
                eb, gc = self.models[name]
                _, z_lik, _ = eb.forward_fused(d[f"{name}_z"], training=True, noise=d[f"{name}_noise_z"])
                ys = d[f"{name}_y"].chunk(NUM_SLICES, 1)
                ss = d[f"{name}_scales"].chunk(NUM_SLICES, 1)
                ms = d[f"{name}_means"].chunk(NUM_SLICES, 1)
                ns = d[f"{name}_noise_y"].chunk(NUM_SLICES, 1)
                liks = []
                for y_s, s_s, m_s, n_s in zip(ys, ss, ms, ns):                     # image_model.py:164-190
                    liks.append(gc.forward_fused(y_s, s_s, m_s, training=True, noise=n_s)[1])
                y_lik = torch.cat(liks, 1)                                          # image_model.py:191
                bpp[name] = (torch.log(y_lik).sum() + torch.log(z_lik).sum()) * scale   # video_model.py:39-42
        for st in self._streams:
            main.wait_stream(st)
        loss = bpp["mv"] + bpp["res"]
        return outs, loss, bpp

    def _backward(self, outs, loss):
        # the distortion gradient reaches the warped tensors as fixed cotangents
        cots = list(self.cot["spynet"]) + [self.cot["warped_frame"], self.cot["warped_feature"]]
        torch.autograd.backward(outs + [loss], cots + [None])

    def step(self):
        """Forward + backward; gradients land in ``.grad`` of the leaves / parameters."""
        for t in self.leaves() + self.params:
            t.grad = None
        outs, loss, bpp = self.forward()
        self._backward(outs, loss)
        self.loss, self.bpp = loss.detach(), {k: v.detach() for k, v in bpp.items()}
        return self.loss

    # ------------------------------------------------------------------ CUDA graph
    def capture(self, warmup: int = 3):
        dev = self.inp["feature"].device
        s = torch.cuda.Stream(dev)
        s.wait_stream(torch.cuda.current_stream(dev))
        with torch.cuda.stream(s):
            for _ in range(warmup):
                self.step()
        torch.cuda.current_stream(dev).wait_stream(s)
        torch.cuda.synchronize(dev)
        for t in self.leaves() + self.params:
            t.grad = None
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g):
            outs, loss, bpp = self.forward()
            self._backward(outs, loss)
            self.loss, self.bpp = loss.detach(), {k: v.detach() for k, v in bpp.items()}
        self._graph = g
        return g

    def replay(self):
        self._graph.replay()

    def grads(self):
        out = {k: self.inp[k].grad for k in LEAVES}
        out["pyr_flow"] = [f.grad for f in self.inp["pyr_flow"]]
        return out
