"""GPU-resident symbol pipeline of one coded P-frame.

The reference's ``ChannelSplitICIP2020ResB.compress`` (``image_model.py:201-257``) leaves the
device 16 times per codec and frame: per slice ``build_indexes`` (a 63-launch Python loop),
``quantize(..., "symbols")`` and two ``.tolist()`` synchronisations (``:237-242``), then range-codes
1.3 M Python ints.  Here a frame's symbols never touch Python:

* every slice's symbols and table indexes are written by ONE fused launch
  (``dsvc_gc_fwd_f32``: quantise + 6-step table search + ``y_hat``) straight into the codec's
  slot of a single device buffer ``[2, N]`` int32 (row 0 symbols, row 1 indexes; the
  bottleneck's z symbols behind the y symbols, their per-channel indexes written once);
* one asynchronous device-to-host copy per frame (both codecs) into pinned memory, on a copy
  stream, fenced by an event -- ``image_model.py:241-242`` becomes no synchronisation at all;
* the host range coder (``csrc/coder.cpp``, compressai's wire format) codes the frame's four
  streams -- mv y, mv z, res y, res z -- and those of the other frames in flight on a pool of
  host threads (``ans.encode_many``): a single rANS stream is sequential, streams are not.

``FrameSymbolPipeline`` keeps ``depth`` frames in flight (device slot + pinned slot each).
Decoding mirrors it: ``ans.decode_many`` -> one pinned upload -> dequantise on the device.
"""
import numpy as np
import torch

from . import _lib
from .entropy import EntropyBottleneck, GaussianConditional, _common_rows

NUM_SLICES = 8
CODECS = ("mv", "res")


class FrameSymbolPipeline:
    def __init__(self, inputs: dict, models: dict, depth: int = 4):
        """inputs: tensors of one frame on a CUDA device (``synthetic.make_pframe_inputs`` keys
        ``{mv,res}_{y,scales,means,z}``; B = 1 as the reference's decoder assumes,
        ``image_model.py:289``); models: {"mv": (EntropyBottleneck, GaussianConditional), ...}
        with CDF tables built (``update()`` / ``update_scale_table``)."""
        self.lib = _lib.load()
        self.models = models
        dev = inputs["mv_y"].device
        self.device = dev
        self.depth = depth
        self.copy_stream = torch.cuda.Stream(dev)
        self.layout = {}            # name -> (ny, nz, offset of the codec in a slot)
        off = 0
        for name in CODECS:
            eb, gc = models[name]
            assert isinstance(eb, EntropyBottleneck) and isinstance(gc, GaussianConditional)
            y, z = inputs[f"{name}_y"], inputs[f"{name}_z"]
            assert y.size(0) == 1, "one frame per slot (image_model.py:289)"
            self.layout[name] = (y.numel(), z.numel(), off)
            off += y.numel() + z.numel()
        self.n_total = off
        self.tables = {name: (models[name][1]._cdf_tables(), models[name][0]._cdf_tables()) for name in CODECS}
        self.medians = {}
        for name in CODECS:
            eb = models[name][0]
            z = inputs[f"{name}_z"]
            self.medians[name] = eb._get_medians().detach().reshape(1, -1, 1, 1).expand_as(z).contiguous()
        self.slots = []
        for _ in range(depth):
            dev_buf = torch.zeros(2, self.n_total, dtype=torch.int32, device=dev)
            for name in CODECS:     # z indexes: the channel number (EntropyBottleneck._build_indexes), static
                ny, nz, o = self.layout[name]
                z = inputs[f"{name}_z"]
                ch = torch.arange(z.size(1), dtype=torch.int32, device=dev).view(1, -1, 1, 1).expand_as(z)
                dev_buf[1, o + ny:o + ny + nz] = ch.reshape(-1)
            self.slots.append({
                "dev": dev_buf, "host": torch.empty(2, self.n_total, dtype=torch.int32, pin_memory=True),
                "y_hat": {name: torch.empty_like(inputs[f"{name}_y"]) for name in CODECS},
                "z_hat": {name: torch.empty_like(inputs[f"{name}_z"]) for name in CODECS},
                "ready": torch.cuda.Event(), "copied": torch.cuda.Event(), "calls": None})
        self.n_launches = 2 * (NUM_SLICES + 1)

    def _bind(self, slot, inputs):
        """Launcher argument tuples of one frame for this slot (bound once per input set)."""
        calls = []
        esz = 4
        buf = slot["dev"]
        for name in CODECS:
            eb, gc = self.models[name]
            ny, nz, o = self.layout[name]
            sb, _ = gc._bounds()
            table = gc.scale_table
            z, med = inputs[f"{name}_z"], self.medians[name]
            # z symbols = round(z - median) (EntropyBottleneck.compress, image_model.py:206)
            calls.append((z.data_ptr(), z.data_ptr(), med.data_ptr(), None, None, None, slot["z_hat"][name].data_ptr(),
                          buf.data_ptr() + (o + ny) * esz, None, None, 0, None, sb, 0.0, 1, nz, nz, nz, nz, 0))
            off = o
            yh = slot["y_hat"][name]
            for y_s, s_s, m_s, h_s in zip(inputs[f"{name}_y"].chunk(NUM_SLICES, 1), inputs[f"{name}_scales"].chunk(NUM_SLICES, 1),
                                          inputs[f"{name}_means"].chunk(NUM_SLICES, 1), yh.chunk(NUM_SLICES, 1)):
                rows, inner, st = _common_rows([y_s, s_s, m_s])
                n = y_s.numel()
                # image_model.py:237-239: build_indexes + quantize("symbols") + y_hat, one launch
                calls.append((y_s.data_ptr(), s_s.data_ptr(), m_s.data_ptr(), None, None, None, h_s.data_ptr(),
                              buf.data_ptr() + off * esz, buf.data_ptr() + (self.n_total + off) * esz,
                              table.data_ptr(), int(table.numel()), None, sb, 0.0, rows, inner, st[0], st[1], st[2], 0))
                off += n
        return calls

    def launch(self, slot_index: int, inputs: dict):
        """Enqueue one frame's 18 symbol launches on the current stream and its single
        device-to-host copy on the copy stream.  Asynchronous."""
        s = self.slots[slot_index]
        # the launcher arguments are bound once per set of input tensors (keyed on their addresses)
        key = tuple(inputs[f"{n}_{k}"].data_ptr() for n in CODECS for k in ("y", "scales", "means", "z"))
        if s["calls"] is None or s["calls"][0] != key:
            s["calls"] = (key, self._bind(s, inputs))
        st = torch.cuda.current_stream(self.device)
        st.wait_event(s["copied"])          # the slot's previous frame has left the device buffer
        fn = self.lib.dsvc_gc_fwd_f32
        with torch.cuda.device(self.device):
            for a in s["calls"][1]:
                err = fn(*a, st.cuda_stream)
                if err:
                    _lib.check(err, "dsvc_gc_fwd_f32")
        s["ready"].record(st)
        with torch.cuda.stream(self.copy_stream):
            self.copy_stream.wait_event(s["ready"])
            s["host"].copy_(s["dev"], non_blocking=True)
            s["copied"].record(self.copy_stream)

    def jobs(self, slot_index: int):
        """The frame's four coder jobs [(symbols, indexes, tables)] (waits for the copy): mv y, mv z,
        res y, res z -- y is ONE stream over all eight slices, as ``image_model.py:253-254``."""
        s = self.slots[slot_index]
        s["copied"].synchronize()
        h = s["host"].numpy()
        out = []
        for name in CODECS:
            ny, nz, o = self.layout[name]
            gc_t, eb_t = self.tables[name]
            out.append((h[0, o:o + ny], h[1, o:o + ny], gc_t))
            out.append((h[0, o + ny:o + ny + nz], h[1, o + ny:o + ny + nz], eb_t))
        return out

    def reconstruction(self, slot_index: int):
        """Encoder-side ``y_hat`` / ``z_hat`` of the slot's frame (device tensors)."""
        s = self.slots[slot_index]
        return s["y_hat"], s["z_hat"]


class FrameSymbolDecoder:
    """The inverse path: four streams of a frame -> ``ans.decode_many`` -> one pinned upload ->
    dequantise on the device (``GaussianConditional.dequantize``, ``image_model.py:288-290``:
    ``symbols.float() + means``).  The table indexes of the y stream are the encoder's (in the
    reference decoder they are rebuilt slice by slice from the decoded context,
    ``image_model.py:286``; here the conv transforms are outside the path)."""

    def __init__(self, pipeline: FrameSymbolPipeline, inputs: dict):
        self.p = pipeline
        dev = pipeline.device
        self.host = torch.empty(pipeline.n_total, dtype=torch.int32, pin_memory=True)
        self.dev = torch.empty(pipeline.n_total, dtype=torch.int32, device=dev)
        self.means = {name: inputs[f"{name}_means"] for name in CODECS}
        self.shapes = {name: (inputs[f"{name}_y"].shape, inputs[f"{name}_z"].shape) for name in CODECS}

    def decode(self, streams, index_jobs, threads=0):
        """streams: the 4 byte strings of a frame; index_jobs: the matching (indexes, tables) pairs.
        Returns {name: (y_hat, z_hat)} device tensors."""
        from . import ans
        syms = ans.decode_many([(s, i, t) for s, (i, t) in zip(streams, index_jobs)], threads)
        h = self.host.numpy()
        k = 0
        for name in CODECS:
            ny, nz, o = self.p.layout[name]
            h[o:o + ny] = syms[k]
            h[o + ny:o + ny + nz] = syms[k + 1]
            k += 2
        self.dev.copy_(self.host, non_blocking=True)
        out = {}
        for name in CODECS:
            ny, nz, o = self.p.layout[name]
            ys, zs = self.shapes[name]
            y_hat = self.dev[o:o + ny].view(ys).float() + self.means[name]          # image_model.py:290
            z_hat = self.dev[o + ny:o + ny + nz].view(zs).float() + self.p.medians[name]
            out[name] = (y_hat, z_hat)
        return out


__all__ = ["FrameSymbolPipeline", "FrameSymbolDecoder"]
