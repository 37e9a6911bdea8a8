// Host-side range coder and CDF-table quantiser ("next" rows f-1 / f-2 of SURVEY.md 8f).
//
// Replaces the native extension of the reference's un-vendored dependency
// (compressai 1.2.1: cpp_exts/rans/rans_interface.cpp over ryg_rans' rans64.h, and
// cpp_exts/ops/ops.cpp::pmf_to_quantized_cdf) as used at
// /root/reference/image_model.py:217-221,253-254,266-274,288 and by update()
// (image_model.py:319-324).  The wire format is the dependency's: rANS with a 64-bit
// state, 32-bit little-endian renormalisation words, lower bound 2^31, 16-bit
// probabilities, and 4-bit bypass chunks for symbols outside the tabulated range -- so
// streams interoperate with compressai's BufferedRansEncoder / RansDecoder.
//
// What is different from the reference pipeline: symbols and table indexes arrive as
// flat int32 buffers copied once from the device (pinned), not as Python lists built
// with .tolist() per slice (image_model.py:241-242).
#include <algorithm>
#include <atomic>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

#include "../../include/deepsvc_b200.h"

namespace {

constexpr int kPrecision = 16;
constexpr int kBypassPrecision = 4;
constexpr uint32_t kMaxBypassVal = (1u << kBypassPrecision) - 1;
constexpr uint64_t kRansL = 1ull << 31;

struct Sym {
    uint16_t start, range;
    bool bypass;
};

struct Encoder {
    std::vector<Sym> syms;
};

constexpr int kLutBits = 8;  // symbol search starts from a 256-entry table per CDF

struct Decoder {
    std::vector<uint32_t> words;
    size_t pos = 0;
    uint64_t x = 0;
    bool ok = false;
    // lut[ci << kLutBits | b] = the symbol whose interval contains the cumulative value
    // b << (16 - kLutBits); rebuilt when the tables change
    std::vector<uint16_t> lut;
    const int32_t* lut_cdfs = nullptr;
    int lut_n = 0, lut_stride = 0;
};

struct Tables {
    const int32_t* cdfs;
    int n_cdfs, stride;
    const int32_t* sizes;
    const int32_t* offsets;
};

inline uint32_t get_bits(Decoder& d, uint32_t n_bits, bool& err) {
    uint64_t x = d.x;
    const uint32_t val = (uint32_t)(x & ((1u << n_bits) - 1));
    x >>= n_bits;
    if (x < kRansL) {
        if (d.pos >= d.words.size()) { err = true; return 0; }
        x = (x << 32) | d.words[d.pos++];
    }
    d.x = x;
    return val;
}

}  // namespace

extern "C" {

int dsvc_pmf_to_quantized_cdf_host(const float* pmf, int n, int precision, int32_t* cdf_out) {
    if (!pmf || !cdf_out || n <= 0 || precision < 1 || precision > 16) return DSVC_ERR_INVALID_ARG;
    for (int i = 0; i < n; ++i)
        if (pmf[i] < 0 || !std::isfinite(pmf[i])) return DSVC_ERR_INVALID_ARG;
    std::vector<uint32_t> cdf(n + 1);
    cdf[0] = 0;
    for (int i = 0; i < n; ++i) cdf[i + 1] = (uint32_t)std::round(pmf[i] * (float)(1 << precision));
    uint32_t total = 0;
    for (uint32_t v : cdf) total += v;
    if (total == 0) return DSVC_ERR_INVALID_ARG;
    for (auto& v : cdf) v = (uint32_t)(((uint64_t)(1u << precision) * v) / total);
    for (int i = 1; i <= n; ++i) cdf[i] += cdf[i - 1];
    cdf[n] = 1u << precision;
    for (int i = 0; i < n; ++i) {
        if (cdf[i] == cdf[i + 1]) {
            // steal one count from the least probable symbol that can spare it
            uint32_t best_freq = ~0u;
            int best = -1;
            for (int j = 0; j < n; ++j) {
                const uint32_t f = cdf[j + 1] - cdf[j];
                if (f > 1 && f < best_freq) { best_freq = f; best = j; }
            }
            if (best < 0) return DSVC_ERR_INVALID_ARG;
            if (best < i) {
                for (int j = best + 1; j <= i; ++j) cdf[j]--;
            } else {
                for (int j = i + 1; j <= best; ++j) cdf[j]++;
            }
        }
    }
    for (int i = 0; i <= n; ++i) cdf_out[i] = (int32_t)cdf[i];
    return 0;
}

void* dsvc_rans_encoder_create(void) { return new Encoder(); }
void dsvc_rans_encoder_destroy(void* h) { delete static_cast<Encoder*>(h); }

int dsvc_rans_encoder_push(void* h, const int32_t* symbols, const int32_t* indexes, int64_t n,
                           const int32_t* cdfs, int n_cdfs, int cdf_stride,
                           const int32_t* cdf_sizes, const int32_t* offsets) {
    if (!h || n < 0 || (n > 0 && (!symbols || !indexes)) || !cdfs || !cdf_sizes || !offsets)
        return DSVC_ERR_INVALID_ARG;
    Encoder& e = *static_cast<Encoder*>(h);
    e.syms.reserve(e.syms.size() + (size_t)n);
    for (int64_t i = 0; i < n; ++i) {
        const int32_t ci = indexes[i];
        if (ci < 0 || ci >= n_cdfs) return DSVC_ERR_INVALID_ARG;
        const int32_t* cdf = cdfs + (size_t)ci * cdf_stride;
        const int32_t max_value = cdf_sizes[ci] - 2;
        if (max_value < 0 || max_value + 1 >= cdf_stride + 0 + 1) return DSVC_ERR_INVALID_ARG;
        int32_t value = symbols[i] - offsets[ci];
        uint32_t raw = 0;
        if (value < 0) {
            raw = (uint32_t)(-2 * (int64_t)value - 1);
            value = max_value;
        } else if (value >= max_value) {
            raw = (uint32_t)(2 * ((int64_t)value - max_value));
            value = max_value;
        }
        e.syms.push_back({(uint16_t)cdf[value], (uint16_t)(cdf[value + 1] - cdf[value]), false});
        if (value == max_value) {
            int32_t n_bypass = 0;
            while (n_bypass < 8 && (raw >> (n_bypass * kBypassPrecision)) != 0) ++n_bypass;
            int32_t val = n_bypass;
            while (val >= (int32_t)kMaxBypassVal) {
                e.syms.push_back({(uint16_t)kMaxBypassVal, (uint16_t)(kMaxBypassVal + 1), true});
                val -= kMaxBypassVal;
            }
            e.syms.push_back({(uint16_t)val, (uint16_t)(val + 1), true});
            for (int32_t j = 0; j < n_bypass; ++j) {
                const uint32_t v = (raw >> (j * kBypassPrecision)) & kMaxBypassVal;
                e.syms.push_back({(uint16_t)v, (uint16_t)(v + 1), true});
            }
        }
    }
    return 0;
}

/* Upper bound of the flushed stream size in bytes for the symbols pushed so far. */
int64_t dsvc_rans_encoder_bound(void* h) {
    if (!h) return 0;
    return (int64_t)(static_cast<Encoder*>(h)->syms.size() + 2) * 4;
}

int dsvc_rans_encoder_flush(void* h, uint8_t* out, int64_t out_cap, int64_t* out_len) {
    if (!h || !out || !out_len) return DSVC_ERR_INVALID_ARG;
    Encoder& e = *static_cast<Encoder*>(h);
    std::vector<uint32_t> buf(e.syms.size() + 2);
    uint32_t* ptr = buf.data() + buf.size();
    uint64_t x = kRansL;
    for (size_t k = e.syms.size(); k-- > 0;) {
        const Sym s = e.syms[k];
        if (!s.bypass) {
            const uint32_t freq = s.range;
            const uint64_t x_max = ((kRansL >> kPrecision) << 32) * freq;
            if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
            x = ((x / freq) << kPrecision) + (x % freq) + s.start;
        } else {
            const uint32_t freq = 1u << (16 - kBypassPrecision);
            const uint64_t x_max = ((kRansL >> 16) << 32) * freq;
            if (x >= x_max) { *--ptr = (uint32_t)x; x >>= 32; }
            x = (x << kBypassPrecision) | s.start;
        }
    }
    ptr -= 2;
    ptr[0] = (uint32_t)x;
    ptr[1] = (uint32_t)(x >> 32);
    const int64_t nbytes = (int64_t)((buf.data() + buf.size()) - ptr) * 4;
    e.syms.clear();
    if (nbytes > out_cap) return DSVC_ERR_INVALID_ARG;
    std::memcpy(out, ptr, (size_t)nbytes);  // little-endian host
    *out_len = nbytes;
    return 0;
}

void* dsvc_rans_decoder_create(const uint8_t* stream, int64_t len) {
    Decoder* d = new Decoder();
    if (stream && len >= 8) {
        d->words.resize((size_t)len / 4);
        std::memcpy(d->words.data(), stream, d->words.size() * 4);
        d->x = (uint64_t)d->words[0] | ((uint64_t)d->words[1] << 32);
        d->pos = 2;
        d->ok = true;
    }
    return d;
}
void dsvc_rans_decoder_destroy(void* h) { delete static_cast<Decoder*>(h); }

int dsvc_rans_decoder_decode(void* h, const int32_t* indexes, int64_t n, const int32_t* cdfs,
                             int n_cdfs, int cdf_stride, const int32_t* cdf_sizes,
                             const int32_t* offsets, int32_t* out) {
    if (!h || n < 0 || (n > 0 && (!indexes || !out)) || !cdfs || !cdf_sizes || !offsets)
        return DSVC_ERR_INVALID_ARG;
    Decoder& d = *static_cast<Decoder*>(h);
    if (!d.ok) return DSVC_ERR_INVALID_ARG;
    const uint32_t mask = (1u << kPrecision) - 1;
    bool err = false;
    if (d.lut_cdfs != cdfs || d.lut_n != n_cdfs || d.lut_stride != cdf_stride) {
        d.lut.assign((size_t)n_cdfs << kLutBits, 0);
        for (int ci = 0; ci < n_cdfs; ++ci) {
            const int32_t* cdf = cdfs + (size_t)ci * cdf_stride;
            const int32_t size = cdf_sizes[ci];
            if (size < 2 || size > cdf_stride) return DSVC_ERR_INVALID_ARG;
            int sidx = 0;
            for (int b = 0; b < (1 << kLutBits); ++b) {
                const uint32_t cum = (uint32_t)b << (kPrecision - kLutBits);
                while (sidx + 2 < size && (uint32_t)cdf[sidx + 1] <= cum) ++sidx;
                d.lut[((size_t)ci << kLutBits) | b] = (uint16_t)sidx;
            }
        }
        d.lut_cdfs = cdfs; d.lut_n = n_cdfs; d.lut_stride = cdf_stride;
    }
    for (int64_t i = 0; i < n; ++i) {
        const int32_t ci = indexes[i];
        if (ci < 0 || ci >= n_cdfs) return DSVC_ERR_INVALID_ARG;
        const int32_t* cdf = cdfs + (size_t)ci * cdf_stride;
        const int32_t size = cdf_sizes[ci];
        const int32_t max_value = size - 2;
        const uint32_t cum = (uint32_t)(d.x & mask);
        // the symbol whose interval [cdf[s], cdf[s+1]) contains cum (the table is strictly
        // increasing): start from the 256-entry table, walk up (a step or two for peaked pmfs)
        int32_t s = d.lut[((size_t)ci << kLutBits) | (cum >> (kPrecision - kLutBits))];
        while (s + 2 < size && (uint32_t)cdf[s + 1] <= cum) ++s;
        if (s < 0 || s + 1 >= size || (uint32_t)cdf[s] > cum || (uint32_t)cdf[s + 1] <= cum) return DSVC_ERR_INVALID_ARG;
        const uint32_t start = (uint32_t)cdf[s], freq = (uint32_t)(cdf[s + 1] - cdf[s]);
        uint64_t x = d.x;
        x = freq * (x >> kPrecision) + (x & mask) - start;
        if (x < kRansL) {
            if (d.pos >= d.words.size()) return DSVC_ERR_INVALID_ARG;
            x = (x << 32) | d.words[d.pos++];
        }
        d.x = x;
        int32_t value = s;
        if (value == max_value) {
            int32_t val = (int32_t)get_bits(d, kBypassPrecision, err);
            int32_t n_bypass = val;
            while (val == (int32_t)kMaxBypassVal && !err) {
                val = (int32_t)get_bits(d, kBypassPrecision, err);
                n_bypass += val;
            }
            uint32_t raw = 0;
            for (int32_t j = 0; j < n_bypass && !err; ++j) {
                val = (int32_t)get_bits(d, kBypassPrecision, err);
                raw |= (uint32_t)val << (j * kBypassPrecision);
            }
            if (err) return DSVC_ERR_INVALID_ARG;
            value = (int32_t)(raw >> 1);
            if (raw & 1) value = -value - 1; else value += max_value;
        }
        out[i] = value + offsets[ci];
    }
    return 0;
}


}  // extern "C"

/* ---- many independent streams at once (one stream = one BufferedRansEncoder's worth: e.g. the y
 * symbols of all 8 slices of one codec of one frame).  A single rANS stream is sequential by
 * construction (one 64-bit state, symbols in reverse), so the parallelism of a byte-compatible
 * coder is ACROSS streams: mv / res, y / z, and the frames in flight.  Streams are handed to
 * `n_threads` host threads through an atomic cursor. */

namespace {

// One stream, encoded straight from the symbol / index arrays in reverse (no intermediate
// symbol vector): emitted words are collected and laid out as [state lo, state hi, words
// in reverse emission order] -- the bytes of dsvc_rans_encoder_push + _flush.
int encode_stream(const int32_t* symbols, const int32_t* indexes, int64_t n, const Tables& t,
                  std::vector<uint32_t>& words, uint8_t* out, int64_t out_cap, int64_t* out_len) {
    words.clear();
    words.reserve((size_t)n / 2 + 16);
    uint64_t x = kRansL;
    auto put = [&](uint32_t start, uint32_t freq) {
        const uint64_t x_max = ((kRansL >> kPrecision) << 32) * freq;
        if (x >= x_max) { words.push_back((uint32_t)x); x >>= 32; }
        x = ((x / freq) << kPrecision) + (x % freq) + start;
    };
    auto put_bits = [&](uint32_t val) {
        const uint32_t freq = 1u << (16 - kBypassPrecision);
        const uint64_t x_max = ((kRansL >> 16) << 32) * freq;
        if (x >= x_max) { words.push_back((uint32_t)x); x >>= 32; }
        x = (x << kBypassPrecision) | val;
    };
    for (int64_t i = n - 1; i >= 0; --i) {
        const int32_t ci = indexes[i];
        if (ci < 0 || ci >= t.n_cdfs) return DSVC_ERR_INVALID_ARG;
        const int32_t* cdf = t.cdfs + (size_t)ci * t.stride;
        const int32_t max_value = t.sizes[ci] - 2;
        if (max_value < 0 || max_value + 1 >= t.stride + 1) return DSVC_ERR_INVALID_ARG;
        int32_t value = symbols[i] - t.offsets[ci];
        uint32_t raw = 0;
        if (value < 0) {
            raw = (uint32_t)(-2 * (int64_t)value - 1);
            value = max_value;
        } else if (value >= max_value) {
            raw = (uint32_t)(2 * ((int64_t)value - max_value));
            value = max_value;
        }
        if (value == max_value) {
            // pushed order: main, count digits (15, 15, ..., rest), raw chunks 0..n-1 -> reversed here
            int32_t n_bypass = 0;
            while (n_bypass < 8 && (raw >> (n_bypass * kBypassPrecision)) != 0) ++n_bypass;
            for (int32_t j = n_bypass - 1; j >= 0; --j) put_bits((raw >> (j * kBypassPrecision)) & kMaxBypassVal);
            int32_t val = n_bypass, n15 = 0;
            while (val >= (int32_t)kMaxBypassVal) { ++n15; val -= kMaxBypassVal; }
            put_bits((uint32_t)val);
            for (int32_t j = 0; j < n15; ++j) put_bits(kMaxBypassVal);
        }
        put((uint32_t)cdf[value], (uint32_t)(cdf[value + 1] - cdf[value]));
    }
    const int64_t nbytes = (int64_t)(words.size() + 2) * 4;
    if (nbytes > out_cap) return DSVC_ERR_INVALID_ARG;
    const uint32_t lo = (uint32_t)x, hi = (uint32_t)(x >> 32);
    std::memcpy(out, &lo, 4);
    std::memcpy(out + 4, &hi, 4);
    for (size_t k = 0; k < words.size(); ++k) std::memcpy(out + 8 + 4 * k, &words[words.size() - 1 - k], 4);
    *out_len = nbytes;
    return 0;
}

template <class F>
int run_streams(int n_streams, int n_threads, F&& one) {
    if (n_streams <= 0) return 0;
    n_threads = std::max(1, std::min(n_threads, n_streams));
    std::atomic<int> next{0}, err{0};
    auto worker = [&]() {
        for (;;) {
            const int s = next.fetch_add(1);
            if (s >= n_streams) break;
            const int e = one(s);
            if (e) err.store(e);
        }
    };
    if (n_threads == 1) {
        worker();
    } else {
        std::vector<std::thread> th;
        for (int k = 0; k < n_threads; ++k) th.emplace_back(worker);
        for (auto& t : th) t.join();
    }
    return err.load();
}

}  // namespace

extern "C" {

int dsvc_rans_encode_many(const int32_t* const* symbols, const int32_t* const* indexes, const int64_t* counts,
                          int n_streams, const int32_t* const* cdfs, const int32_t* n_cdfs,
                          const int32_t* cdf_strides, const int32_t* const* cdf_sizes,
                          const int32_t* const* offsets, uint8_t* const* out, const int64_t* out_cap,
                          int64_t* out_len, int n_threads) {
    if (n_streams < 0 || (n_streams > 0 && (!symbols || !indexes || !counts || !cdfs || !n_cdfs || !cdf_strides ||
                                            !cdf_sizes || !offsets || !out || !out_cap || !out_len)))
        return DSVC_ERR_INVALID_ARG;
    return run_streams(n_streams, n_threads, [&](int s) -> int {
        if (counts[s] < 0 || (counts[s] > 0 && (!symbols[s] || !indexes[s])) || !out[s]) return DSVC_ERR_INVALID_ARG;
        thread_local std::vector<uint32_t> words;
        const Tables t{cdfs[s], n_cdfs[s], cdf_strides[s], cdf_sizes[s], offsets[s]};
        return encode_stream(symbols[s], indexes[s], counts[s], t, words, out[s], out_cap[s], &out_len[s]);
    });
}

int dsvc_rans_decode_many(const uint8_t* const* streams, const int64_t* stream_len, const int32_t* const* indexes,
                          const int64_t* counts, int n_streams, const int32_t* const* cdfs, const int32_t* n_cdfs,
                          const int32_t* cdf_strides, const int32_t* const* cdf_sizes,
                          const int32_t* const* offsets, int32_t* const* out, int n_threads) {
    if (n_streams < 0 || (n_streams > 0 && (!streams || !stream_len || !indexes || !counts || !cdfs || !n_cdfs ||
                                            !cdf_strides || !cdf_sizes || !offsets || !out)))
        return DSVC_ERR_INVALID_ARG;
    return run_streams(n_streams, n_threads, [&](int s) -> int {
        void* h = dsvc_rans_decoder_create(streams[s], stream_len[s]);
        const int e = dsvc_rans_decoder_decode(h, indexes[s], counts[s], cdfs[s], n_cdfs[s], cdf_strides[s],
                                               cdf_sizes[s], offsets[s], out[s]);
        dsvc_rans_decoder_destroy(h);
        return e;
    });
}

}  // extern "C"
