"""Host-side range coder / CDF quantiser (rows f-1, f-2) against the oracle's pure-Python
restatement of compressai's native extension: byte-identical streams, identical tables."""
import numpy as np
import pytest
import torch


def _tables(oracle, ch=4, seed=5):
    eb, gc = oracle.make_entropy_models(ch, seed=seed)
    return eb, gc


def test_pmf_to_quantized_cdf_matches_oracle(oracle):
    from compressai import ans as oans
    from deepsvc_b200 import ans
    rng = np.random.default_rng(0)
    for n in (2, 3, 17, 200, 1500):
        pmf = rng.random(n).astype(np.float32) ** 6       # many tiny entries -> stealing loop
        pmf /= pmf.sum()
        assert ans.pmf_to_quantized_cdf(pmf.tolist()) == oans.pmf_to_quantized_cdf(pmf.tolist())
    with pytest.raises(ValueError):
        ans.pmf_to_quantized_cdf([0.5, -0.1])
    with pytest.raises(ValueError):
        ans.pmf_to_quantized_cdf([0.0, 0.0])


def test_cdf_tables_match_oracle_update(oracle):
    import deepsvc_b200 as d
    from deepsvc_b200 import cdf
    eb_o, gc_o = _tables(oracle, 8, seed=1)
    c, o, l = cdf.gaussian_cdf_tables(gc_o.scale_table, 1e-9)
    assert torch.equal(c, gc_o.quantized_cdf) and torch.equal(o, gc_o.offset) and torch.equal(l, gc_o.cdf_length)
    eb = d.EntropyBottleneck(8)
    eb.load_state_dict(eb_o.state_dict(), strict=False)
    eb_o.update(force=True)
    assert eb.update(force=True) is True and eb.update() is False
    assert torch.equal(eb.quantized_cdf, eb_o.quantized_cdf)
    assert torch.equal(eb.offset, eb_o.offset) and torch.equal(eb.cdf_length, eb_o.cdf_length)
    gc = d.GaussianConditional(None)
    assert gc.update_scale_table(oracle.get_scale_table()) is True
    assert torch.equal(gc.quantized_cdf, gc_o.quantized_cdf)
    assert gc.update_scale_table(oracle.get_scale_table()) is False   # already built, not forced


def test_rans_bytes_identical_and_roundtrip(oracle):
    from compressai import ans as oans
    from deepsvc_b200 import ans
    _, gc = _tables(oracle)
    cdf, lens, offs = gc.quantized_cdf, gc.cdf_length, gc.offset
    rng = np.random.default_rng(1)
    n = 6000
    idx = rng.integers(0, 64, size=n).astype(np.int32)
    sym = np.round(rng.normal(0, np.maximum(0.3, 0.11 * 1.13 ** idx))).astype(np.int32)
    sym[10], sym[11], sym[12], sym[13] = 5000, -5000, 70000, -1   # bypass-coded outliers
    ref = oans.BufferedRansEncoder()
    ref.encode_with_indexes(sym.tolist(), idx.tolist(), cdf.tolist(), lens.tolist(), offs.tolist())
    want = ref.flush()
    # array interface, two pushes + one flush (image_model.py:253-254 pattern)
    t = ans.CdfTables(cdf, lens, offs)
    enc = ans.BufferedRansEncoder()
    enc.encode_with_indexes(sym[:2500], idx[:2500], t)
    enc.encode_with_indexes(sym[2500:], idx[2500:], t)
    got = enc.flush()
    assert got == want
    # list interface (drop-in signature)
    enc2 = ans.BufferedRansEncoder()
    enc2.encode_with_indexes(sym.tolist(), idx.tolist(), cdf.tolist(), lens.tolist(), offs.tolist())
    assert enc2.flush() == want
    # slice-by-slice decoding from one stream (image_model.py:273-288 pattern)
    dec = ans.RansDecoder()
    dec.set_stream(want)
    a = dec.decode_stream(idx[:1000].tolist(), cdf.tolist(), lens.tolist(), offs.tolist())
    b = dec.decode_stream_array(idx[1000:], t)
    assert a == sym[:1000].tolist() and np.array_equal(b, sym[1000:])
    # the oracle's decoder reads our stream too
    odec = oans.RansDecoder()
    odec.set_stream(got)
    assert odec.decode_stream(idx.tolist(), cdf.tolist(), lens.tolist(), offs.tolist()) == sym.tolist()
    # empty run and corrupt stream
    e = ans.BufferedRansEncoder()
    assert len(e.flush()) == 8
    bad = ans.RansDecoder()
    bad.set_stream(want[:16])
    with pytest.raises(ValueError):
        bad.decode_stream_array(idx, t)
    with pytest.raises(ValueError):
        ans.BufferedRansEncoder().encode_with_indexes([0], [99], t)


def test_patch_reference_rebinds_names(reference_modules):
    """Drop-in installation into the unmodified reference (build container only)."""
    import deepsvc_b200 as d
    modules, image_model, video_model = reference_modules
    orig = modules.torch_warp
    m = video_model.DeepSVC()                           # built before patching: oracle-shim classes
    try:
        d.patch_reference()
        assert modules.torch_warp is d.torch_warp and video_model.torch_warp is d.torch_warp
        assert image_model.ste_round is d.ste_round
        assert image_model.BufferedRansEncoder is d.ans.BufferedRansEncoder
        assert d.swap_entropy_models(m) == 4
        assert isinstance(m.mv_codec.gaussian_conditional, d.GaussianConditional)
        assert isinstance(m.res_codec.entropy_bottleneck, d.EntropyBottleneck)
        # parameters are shared objects, state_dict keys unchanged
        ref_keys = sorted(video_model.DeepSVC().state_dict().keys())
        assert sorted(m.state_dict().keys()) == ref_keys
        assert m.update(force=True) is True            # builds CDF tables through the C++ quantiser
        assert m.mv_codec.gaussian_conditional.quantized_cdf.shape[0] == 64
        assert float(m.aux_loss()) > 0                 # isinstance(m, EntropyBottleneck) still matches
        m2 = video_model.DeepSVC()                      # built after patching: drop-ins directly
        assert isinstance(m2.mv_codec.entropy_bottleneck, d.EntropyBottleneck)
        assert d.swap_entropy_models(m2) == 0
    finally:
        d.unpatch_reference()
    assert modules.torch_warp is orig


def _rans64_reference(syms, scale_bits=16):
    """rANS with a 64-bit state and 32-bit renormalisation, written out from the published
    recurrences (Duda 2013; F. Giesen's rans64.h, public domain) on Python ints -- independent of
    csrc/coder.cpp and of the oracle's coder: x starts at L = 2^31; symbols are encoded LAST to
    FIRST; encoding (start, freq) first emits the low 32 bits of x while
    x >= ((L >> scale_bits) << 32) * freq, then x <- (x // freq << scale_bits) + x % freq + start;
    the final x follows as two words; the word emitted last comes first in the stream."""
    L = 1 << 31
    x, words = L, []
    for start, freq, bits in reversed(syms):
        if x >= ((L >> bits) << 32) * freq:
            words.append(x & 0xFFFFFFFF)
            x >>= 32
        x = ((x // freq) << bits) + (x % freq) + start
    out = [x & 0xFFFFFFFF, x >> 32] + words[::-1]
    return b"".join(int(w).to_bytes(4, "little") for w in out)


def test_rans64_known_answer_stream():
    """Known-answer test of the wire format: a hand-checkable stream, and a longer one against the
    big-int recurrences above.  Both the C++ coder and the oracle's pure-Python coder must
    reproduce the bytes; compressai's bypass convention (4-bit chunks, count first) is spelled out."""
    from compressai import ans as oans
    from deepsvc_b200 import ans
    # one CDF over 4 symbols with offset 0: freqs 8192, 16384, 32768 and the escape 8192 (2^16 total)
    cdf = [[0, 8192, 24576, 57344, 65536]]
    lens, offs = [5], [0]
    t = ans.CdfTables(cdf, lens, offs)
    # (1) the single symbol 1: x = ((2^31 // 16384) << 16) + 2^31 % 16384 + 8192 = 2^33 + 8192
    #     = 0x2_0000_2000 -> words [0x00002000, 0x00000002], no renormalisation word
    want = bytes.fromhex("00200000" "02000000")
    assert _rans64_reference([(8192, 16384, 16)]) == want
    assert ans.encode_many([([1], [0], t)], 1) == [want]
    e = ans.BufferedRansEncoder()
    e.encode_with_indexes([1], [0], t)
    assert e.flush() == want
    o = oans.BufferedRansEncoder()
    o.encode_with_indexes([1], [0], cdf, lens, offs)
    assert o.flush() == want
    # (2) a longer message incl. out-of-range values.  Symbol value v with max_value = 3 (the last
    # table entry is the escape): v < 0 -> raw = -2 v - 1, v >= 3 -> raw = 2 (v - 3); the escape is
    # followed by the number n of non-zero 4-bit chunks of raw (as digits 15, 15, ..., rest, each a
    # uniform 4-bit symbol) and then the n chunks, least significant first.
    rng = np.random.default_rng(7)
    msg = rng.integers(0, 3, size=400).tolist() + [3, 7, -1, -40, 100000, 0, 2]
    syms = []
    for v in msg:
        raw = 0
        if v < 0:
            raw, v = -2 * v - 1, 3
        elif v >= 3:
            raw, v = 2 * (v - 3), 3
        syms.append((cdf[0][v], cdf[0][v + 1] - cdf[0][v], 16))
        if v == 3:
            n = 0
            while (raw >> (4 * n)) != 0:
                n += 1
            val = n
            while val >= 15:
                syms.append((15, 1, 4))
                val -= 15
            syms.append((val, 1, 4))
            for j in range(n):
                syms.append(((raw >> (4 * j)) & 15, 1, 4))
    want = _rans64_reference(syms)
    idx = [0] * len(msg)
    assert ans.encode_many([(msg, idx, t)], 1) == [want]
    e = ans.BufferedRansEncoder()
    e.encode_with_indexes(msg, idx, t)
    assert e.flush() == want
    o = oans.BufferedRansEncoder()
    o.encode_with_indexes(msg, idx, cdf, lens, offs)
    assert o.flush() == want
    assert ans.decode_many([(want, idx, t)], 1)[0].tolist() == msg
    d = ans.RansDecoder()
    d.set_stream(want)
    assert d.decode_stream(idx, t) == msg
    # streams are independent of how many threads code them
    jobs = [(msg, idx, t)] * 5
    assert ans.encode_many(jobs, 4) == [want] * 5
