"""Range coder with the ``compressai.ans`` interface (``image_model.py:8``):
``BufferedRansEncoder``, ``RansEncoder``, ``RansDecoder``.  The coding loops are the
C++ host functions of ``csrc/coder.cpp``; symbols / indexes / tables are passed as flat
int32 buffers (numpy arrays or CPU tensors -- e.g. one pinned device-to-host copy per
slice) instead of Python lists.  Lists are accepted for signature compatibility.
"""
import ctypes

import numpy as np

from . import _lib


def _i32(a):
    """Contiguous int32 numpy view/copy of a list / numpy array / CPU tensor."""
    if hasattr(a, "detach"):
        a = a.detach()
        if a.is_cuda:
            a = a.cpu()
        a = a.numpy()
    return np.ascontiguousarray(np.asarray(a, dtype=np.int32).reshape(-1))


class CdfTables:
    """Flat int32 views of (quantized_cdf [n, stride], cdf_length [n], offset [n])."""

    def __init__(self, cdfs, cdf_sizes, offsets):
        c = cdfs.detach().cpu().numpy() if hasattr(cdfs, "detach") else np.asarray(cdfs)
        if c.dtype == object or c.ndim != 2:  # ragged python lists
            rows = [list(r) for r in cdfs]
            width = max(len(r) for r in rows)
            c = np.zeros((len(rows), width), dtype=np.int32)
            for i, r in enumerate(rows):
                c[i, :len(r)] = r
        self.cdfs = np.ascontiguousarray(c.astype(np.int32, copy=False))
        self.sizes = _i32(cdf_sizes)
        self.offsets = _i32(offsets)
        self.n, self.stride = self.cdfs.shape
        if not (len(self.sizes) == len(self.offsets) == self.n):
            raise ValueError("cdf tables, sizes and offsets disagree in length")
        if self.n and int(self.sizes.max()) > self.stride:
            raise ValueError("cdf_length exceeds the table width")


def _tables(cdfs, cdf_sizes, offsets):
    return cdfs if isinstance(cdfs, CdfTables) else CdfTables(cdfs, cdf_sizes, offsets)


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


class BufferedRansEncoder:
    def __init__(self):
        self._lib = _lib.load()
        self._h = self._lib.dsvc_rans_encoder_create()

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.dsvc_rans_encoder_destroy(self._h)
            self._h = None

    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes=None, offsets=None):
        t = _tables(cdfs, cdfs_sizes, offsets)
        s, i = _i32(symbols), _i32(indexes)
        if s.size != i.size:
            raise ValueError("symbols and indexes differ in length")
        err = self._lib.dsvc_rans_encoder_push(self._h, _p(s), _p(i), s.size, _p(t.cdfs), t.n,
                                               t.stride, _p(t.sizes), _p(t.offsets))
        if err:
            raise ValueError("rans encoder: index or symbol table out of range")

    def flush(self) -> bytes:
        cap = int(self._lib.dsvc_rans_encoder_bound(self._h))
        buf = np.empty(cap, dtype=np.uint8)
        n = ctypes.c_int64(0)
        err = self._lib.dsvc_rans_encoder_flush(self._h, _p(buf), cap, ctypes.byref(n))
        if err:
            raise RuntimeError("rans encoder flush failed")
        return buf[: n.value].tobytes()


class RansEncoder:
    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes=None, offsets=None) -> bytes:
        e = BufferedRansEncoder()
        e.encode_with_indexes(symbols, indexes, cdfs, cdfs_sizes, offsets)
        return e.flush()


class RansDecoder:
    def __init__(self):
        self._lib = _lib.load()
        self._h = None

    def __del__(self):
        if getattr(self, "_h", None):
            self._lib.dsvc_rans_decoder_destroy(self._h)
            self._h = None

    def set_stream(self, encoded: bytes):
        if self._h:
            self._lib.dsvc_rans_decoder_destroy(self._h)
        buf = np.frombuffer(encoded, dtype=np.uint8)
        self._h = self._lib.dsvc_rans_decoder_create(_p(buf), buf.size)

    def decode_stream_array(self, indexes, cdfs, cdfs_sizes=None, offsets=None) -> np.ndarray:
        """Decoded symbols as an int32 numpy array (no Python list)."""
        if not self._h:
            raise RuntimeError("set_stream() first")
        t = _tables(cdfs, cdfs_sizes, offsets)
        i = _i32(indexes)
        out = np.empty(i.size, dtype=np.int32)
        err = self._lib.dsvc_rans_decoder_decode(self._h, _p(i), i.size, _p(t.cdfs), t.n, t.stride,
                                                 _p(t.sizes), _p(t.offsets), _p(out))
        if err:
            raise ValueError("rans decoder: corrupt stream or tables out of range")
        return out

    def decode_stream(self, indexes, cdfs, cdfs_sizes=None, offsets=None):
        return self.decode_stream_array(indexes, cdfs, cdfs_sizes, offsets).tolist()

    def decode_with_indexes(self, encoded, indexes, cdfs, cdfs_sizes=None, offsets=None):
        self.set_stream(encoded)
        return self.decode_stream(indexes, cdfs, cdfs_sizes, offsets)


def _ptr_array(arrs):
    return (ctypes.c_void_p * len(arrs))(*[a.ctypes.data for a in arrs])


def _table_arrays(tables):
    n = len(tables)
    return (_ptr_array([t.cdfs for t in tables]), (ctypes.c_int32 * n)(*[t.n for t in tables]),
            (ctypes.c_int32 * n)(*[t.stride for t in tables]), _ptr_array([t.sizes for t in tables]),
            _ptr_array([t.offsets for t in tables]))


def encode_many(jobs, threads: int = 0):
    """Encode independent streams in parallel on host threads (``dsvc_rans_encode_many``).
    jobs: [(symbols int32 array, indexes int32 array, CdfTables), ...]; stream k is
    byte-identical to ``BufferedRansEncoder().encode_with_indexes(*jobs[k]); flush()``.
    Returns a list of ``bytes``."""
    import os
    lib = _lib.load()
    n = len(jobs)
    if n == 0:
        return []
    syms = [_i32(j[0]) for j in jobs]
    idxs = [_i32(j[1]) for j in jobs]
    tabs = [j[2] for j in jobs]
    for s, i in zip(syms, idxs):
        if s.size != i.size:
            raise ValueError("symbols and indexes differ in length")
    counts = (ctypes.c_int64 * n)(*[s.size for s in syms])
    cdfs, n_cdfs, strides, sizes, offs = _table_arrays(tabs)
    extra = 0
    while True:
        caps = [(s.size * (2 + 4 * extra) + 1024) * 4 for s in syms]   # 32 bits per symbol is ample for real streams
        outs = [np.empty(c, dtype=np.uint8) for c in caps]
        out_len = (ctypes.c_int64 * n)()
        err = lib.dsvc_rans_encode_many(_ptr_array(syms), _ptr_array(idxs), counts, n, cdfs, n_cdfs, strides, sizes,
                                        offs, _ptr_array(outs), (ctypes.c_int64 * n)(*caps), out_len,
                                        int(threads) or min(n, os.cpu_count() or 1))
        if err == 0:
            return [o[:l].tobytes() for o, l in zip(outs, out_len)]
        if extra >= 2:
            raise ValueError("rans encoder: index or symbol table out of range")
        extra += 1                      # (or every symbol bypass-coded: retry with room for 11 words each)


def decode_many(jobs, threads: int = 0):
    """Decode independent streams in parallel (``dsvc_rans_decode_many``).
    jobs: [(stream bytes, indexes int32 array, CdfTables), ...] -> list of int32 numpy arrays."""
    import os
    lib = _lib.load()
    n = len(jobs)
    if n == 0:
        return []
    bufs = [np.frombuffer(j[0], dtype=np.uint8) for j in jobs]
    idxs = [_i32(j[1]) for j in jobs]
    tabs = [j[2] for j in jobs]
    outs = [np.empty(i.size, dtype=np.int32) for i in idxs]
    cdfs, n_cdfs, strides, sizes, offs = _table_arrays(tabs)
    err = lib.dsvc_rans_decode_many(_ptr_array(bufs), (ctypes.c_int64 * n)(*[b.size for b in bufs]), _ptr_array(idxs),
                                    (ctypes.c_int64 * n)(*[i.size for i in idxs]), n, cdfs, n_cdfs, strides, sizes, offs,
                                    _ptr_array(outs), int(threads) or min(n, os.cpu_count() or 1))
    if err:
        raise ValueError("rans decoder: corrupt stream or tables out of range")
    return outs


def pmf_to_quantized_cdf(pmf, precision: int = 16):
    """``compressai._CXX.pmf_to_quantized_cdf`` (list of float -> list of int)."""
    p = np.ascontiguousarray(np.asarray(pmf, dtype=np.float32).reshape(-1))
    out = np.empty(p.size + 1, dtype=np.int32)
    err = _lib.load().dsvc_pmf_to_quantized_cdf_host(_p(p), p.size, int(precision), _p(out))
    if err:
        raise ValueError("Invalid `pmf`: negative, non-finite or all-zero")
    return out.tolist()


__all__ = ["BufferedRansEncoder", "RansEncoder", "RansDecoder", "CdfTables", "pmf_to_quantized_cdf",
           "encode_many", "decode_many"]
