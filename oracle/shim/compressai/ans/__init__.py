"""Oracle restatement (pure Python) of compressai 1.2.1's native range coder.

ORACLE / TEST INFRASTRUCTURE ONLY; PARITY UNPINNED (no compressai binary or
bitstream fixture exists in the reference or in this image).

Restates ``compressai/cpp_exts/rans/rans_interface.cpp`` (which wraps the public
domain ``ryg_rans`` ``rans64.h``: 64-bit state, 32-bit renormalisation words,
lower bound L = 2**31) and ``compressai/cpp_exts/ops/ops.cpp``
(``pmf_to_quantized_cdf``).  Constants: 16-bit probability precision, 4-bit
bypass chunks for out-of-range symbols.  Reference call sites:
``image_model.py:8,217-221,253-254,266-274,288``.
"""
import math
import struct
from typing import List

PRECISION = 16
BYPASS_PRECISION = 4
MAX_BYPASS_VAL = (1 << BYPASS_PRECISION) - 1
RANS64_L = 1 << 31
_M32 = 0xFFFFFFFF
_M64 = 0xFFFFFFFFFFFFFFFF


def _c_round(x: float) -> int:
    """std::round: half away from zero."""
    return int(math.floor(x + 0.5)) if x >= 0 else -int(math.floor(-x + 0.5))


def pmf_to_quantized_cdf(pmf: List[float], precision: int = 16) -> List[int]:
    """ops.cpp ``pmf_to_quantized_cdf``: float pmf -> strictly increasing integer cdf."""
    import numpy as np

    for p in pmf:
        if p < 0 or not math.isfinite(p):
            raise ValueError(f"Invalid `pmf`, non-finite or negative element found: {p}")
    # the C++ code rounds float(p) * (1 << precision) evaluated in fp32
    cdf = [0] * (len(pmf) + 1)
    scale = np.float32(1 << precision)
    for i, p in enumerate(pmf):
        cdf[i + 1] = _c_round(float(np.float32(p) * scale))
    total = sum(cdf)
    if total == 0:
        raise ValueError("Invalid `pmf`: at least one element must have a non-zero probability.")
    cdf = [((1 << precision) * c) // total for c in cdf]
    acc = 0
    for i in range(len(cdf)):
        acc += cdf[i]
        cdf[i] = acc
    cdf[-1] = 1 << precision
    n = len(cdf)
    for i in range(n - 1):
        if cdf[i] == cdf[i + 1]:
            best_freq = 1 << 62
            best_steal = -1
            for j in range(n - 1):
                freq = cdf[j + 1] - cdf[j]
                if 1 < freq < best_freq:
                    best_freq = freq
                    best_steal = j
            assert best_steal != -1
            if best_steal < i:
                for j in range(best_steal + 1, i + 1):
                    cdf[j] -= 1
            else:
                assert best_steal > i
                for j in range(i + 1, best_steal + 1):
                    cdf[j] += 1
    assert cdf[0] == 0 and cdf[-1] == (1 << precision)
    for i in range(n - 1):
        assert cdf[i + 1] > cdf[i]
    return cdf


class _Sym:
    __slots__ = ("start", "range", "bypass")

    def __init__(self, start, rng, bypass):
        self.start = start & 0xFFFF
        self.range = rng & 0xFFFF
        self.bypass = bypass


def _push_symbols(syms, symbols, indexes, cdfs, cdfs_sizes, offsets):
    for i in range(len(symbols)):
        cdf_idx = indexes[i]
        cdf = cdfs[cdf_idx]
        max_value = cdfs_sizes[cdf_idx] - 2
        value = symbols[i] - offsets[cdf_idx]
        raw_val = 0
        if value < 0:
            raw_val = -2 * value - 1
            value = max_value
        elif value >= max_value:
            raw_val = 2 * (value - max_value)
            value = max_value
        syms.append(_Sym(cdf[value], cdf[value + 1] - cdf[value], False))
        if value == max_value:
            n_bypass = 0
            while (raw_val >> (n_bypass * BYPASS_PRECISION)) != 0:
                n_bypass += 1
            val = n_bypass
            while val >= MAX_BYPASS_VAL:
                syms.append(_Sym(MAX_BYPASS_VAL, MAX_BYPASS_VAL + 1, True))
                val -= MAX_BYPASS_VAL
            syms.append(_Sym(val, val + 1, True))
            for j in range(n_bypass):
                v = (raw_val >> (j * BYPASS_PRECISION)) & MAX_BYPASS_VAL
                syms.append(_Sym(v, v + 1, True))


def _flush(syms) -> bytes:
    x = RANS64_L
    words = []  # emitted in reverse (the C code writes backwards from the buffer end)
    for sym in reversed(syms):
        if not sym.bypass:
            freq = sym.range
            x_max = ((RANS64_L >> PRECISION) << 32) * freq
            if x >= x_max:
                words.append(x & _M32)
                x >>= 32
            x = ((x // freq) << PRECISION) + (x % freq) + sym.start
        else:
            freq = 1 << (16 - BYPASS_PRECISION)
            x_max = ((RANS64_L >> 16) << 32) * freq
            if x >= x_max:
                words.append(x & _M32)
                x >>= 32
            x = ((x << BYPASS_PRECISION) | sym.start) & _M64
    # Rans64EncFlush: low word first
    words.append((x >> 32) & _M32)
    words.append(x & _M32)
    words.reverse()
    return struct.pack(f"<{len(words)}I", *words)


class BufferedRansEncoder:
    def __init__(self):
        self._syms = []

    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets):
        _push_symbols(self._syms, symbols, indexes, cdfs, cdfs_sizes, offsets)

    def flush(self) -> bytes:
        out = _flush(self._syms)
        self._syms = []
        return out


class RansEncoder:
    def encode_with_indexes(self, symbols, indexes, cdfs, cdfs_sizes, offsets) -> bytes:
        enc = BufferedRansEncoder()
        enc.encode_with_indexes(symbols, indexes, cdfs, cdfs_sizes, offsets)
        return enc.flush()


class RansDecoder:
    def __init__(self):
        self._words = ()
        self._pos = 0
        self._x = 0

    def set_stream(self, encoded: bytes):
        n = len(encoded) // 4
        self._words = struct.unpack(f"<{n}I", encoded[: 4 * n])
        self._x = self._words[0] | (self._words[1] << 32)
        self._pos = 2

    def _get_bits(self, n_bits):
        x = self._x
        val = x & ((1 << n_bits) - 1)
        x >>= n_bits
        if x < RANS64_L:
            x = (x << 32) | self._words[self._pos]
            self._pos += 1
        self._x = x
        return val

    def decode_stream(self, indexes, cdfs, cdfs_sizes, offsets):
        out = [0] * len(indexes)
        mask = (1 << PRECISION) - 1
        for i in range(len(indexes)):
            cdf_idx = indexes[i]
            cdf = cdfs[cdf_idx]
            max_value = cdfs_sizes[cdf_idx] - 2
            offset = offsets[cdf_idx]
            cum_freq = self._x & mask
            s = 0
            size = cdfs_sizes[cdf_idx]
            while s < size and not (cdf[s] > cum_freq):
                s += 1
            s -= 1
            start = cdf[s]
            freq = cdf[s + 1] - cdf[s]
            x = self._x
            x = freq * (x >> PRECISION) + (x & mask) - start
            if x < RANS64_L:
                x = (x << 32) | self._words[self._pos]
                self._pos += 1
            self._x = x
            value = s
            if value == max_value:
                val = self._get_bits(BYPASS_PRECISION)
                n_bypass = val
                while val == MAX_BYPASS_VAL:
                    val = self._get_bits(BYPASS_PRECISION)
                    n_bypass += val
                raw_val = 0
                for j in range(n_bypass):
                    val = self._get_bits(BYPASS_PRECISION)
                    raw_val |= val << (j * BYPASS_PRECISION)
                value = raw_val >> 1
                if raw_val & 1:
                    value = -value - 1
                else:
                    value += max_value
            out[i] = value + offset
        return out

    def decode_with_indexes(self, encoded, indexes, cdfs, cdfs_sizes, offsets):
        self.set_stream(encoded)
        return self.decode_stream(indexes, cdfs, cdfs_sizes, offsets)


__all__ = ["BufferedRansEncoder", "RansEncoder", "RansDecoder", "pmf_to_quantized_cdf"]
