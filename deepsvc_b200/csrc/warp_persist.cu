// Backward bilinear warp, persistent shared-memory / TMA staged forward (NCHW fp32, sm_100a).
//
// Replaces /root/reference/modules.py:25-62 for the bandwidth-critical calls (the 64-ch
// feature warp of modules.py:429 is 86 % of the hot path's bytes).
//
// One persistent CTA per resident slot (2 per SM) pulls work units from a device-side
// counter.  A unit is a TW x TH tile of output pixels of one batch item and a channel
// range: whole tiles first, then the last one-wave's worth of tiles cut into `tail_split` (2)
// channel ranges, so that all CTAs run dry within a fraction of a tile time (a plain grid
// of 1020 tiles is 3.45 waves of 296 slots: 14 % of the machine idles in the last wave).
//
// Three roles per CTA, decoupled by mbarrier rings:
//   scout warp    claims the next unit, reads the tile's flow, computes every pixel's source
//                 coordinate (reference arithmetic, warp_common.cuh) and the tile's source
//                 bounding box, and posts a unit descriptor -- one unit ahead of the data;
//   issuer warp   (one elected lane) streams the bounding box, CC channel planes at a time,
//                 into a STAGES-deep shared-memory ring with 3-D TMA loads over the tensor
//                 viewed as (x, plane, y): box = BW x CC planes x 8 rows.  The ring runs on
//                 across unit boundaries, so loads never stop while consumers set up a tile;
//   8 consumer    compute their pixels' taps once per unit and gather the four taps of each
//   warps         pixel from shared memory for every staged channel (row pitch = CC*BW floats
//                 = a multiple of 32 banks: lanes on different source rows do not collide),
//                 storing coalesced 128-byte rows.
//
// A tile whose bounding box does not fit the staging box (wild flow, motion boundaries) is
// re-posted by the scout as four 32 x 16 quadrants, each staged on its own smaller bounding
// box if that fits, else gathered straight from global memory by the consumers (same
// arithmetic, bit-identical results).  Everything is local to the CTA: no work list, no
// follow-up launch.  The last scout to find the counter exhausted re-zeroes it, so the
// 16-byte scheduler state is zero before and after every launch (no memset).
#include <climits>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <vector>

#include "tma_utils.cuh"
#include "warp_common.cuh"

namespace dsvc {

template <int TW_, int TH_, int BW_, int BHMAX_, int CC_, int STAGES_, int MINB_ = 2, int MAXREG_ = 96>
struct PersistCfg {
    // registers per thread: 96 = what __launch_bounds__(320, 2) gave; 72 leaves 19 K registers per SM
    // free, so that the frame's short launches (48-reg few-channel warps, entropy kernels) can be
    // resident next to the two persistent CTAs instead of queueing behind them
    static constexpr int MAXREG = MAXREG_;
    static constexpr int TW = TW_, TH = TH_, BW = BW_, BHMAX = BHMAX_, CC = CC_, STAGES = STAGES_;
    static constexpr int MINB = MINB_;  // resident CTAs per SM the kernel is compiled for
    static constexpr int CONSUMER_WARPS = 8;
    static constexpr int ISSUER_WARP = CONSUMER_WARPS, SCOUT_WARP = CONSUMER_WARPS + 1;
    static constexpr int THREADS = (CONSUMER_WARPS + 2) * 32;
    static constexpr int ROWS_PER_WARP = TH / CONSUMER_WARPS;  // rows of the tile per warp
    static constexpr int XH = TW / 32;                         // 32-pixel column groups
    static constexpr int PPT = ROWS_PER_WARP * XH;             // pixels per consumer thread
    static constexpr int ROWCHUNK = 8;                         // rows per TMA box
    static constexpr int ROW_PITCH = CC * BW;                  // smem floats between box rows
    static constexpr int CHUNK_FLOATS = CC * ROWCHUNK * BW;    // one TMA box [8 rows][CC][BW]
    static constexpr int STAGE_FLOATS = (BHMAX / ROWCHUNK) * CHUNK_FLOATS;
    static constexpr size_t SMEM_BYTES = (size_t)STAGES * STAGE_FLOATS * sizeof(float);
    static constexpr int QY_ROWS_DIV = TH > 16 ? 2 : 1;
    static constexpr int NDESC = 2;  // unit descriptors in flight (scout runs ahead)
    static constexpr int SCOUT_RB = TH / QY_ROWS_DIV;  // scout: rows per batch (flow loads in flight)
    static_assert(TH % CONSUMER_WARPS == 0 && TW % 32 == 0 && BHMAX % ROWCHUNK == 0, "tile shape");
    static constexpr int QX = TW >= 64 ? 2 : 1, QY = 2;  // sub-rectangles of an unstageable tile
    static_assert((TW & (TW - 1)) == 0 && TH % QY == 0 && (TW / QX) % 32 == 0, "sub-rectangles are whole lane groups");
    static_assert((BW * 4) % 16 == 0 && (CHUNK_FLOATS * 4) % 128 == 0, "TMA alignment");
    static_assert(ROW_PITCH % 16 == 0, "row pitch: a multiple of the 32 banks (or of 16: slightly more conflicts)");
};

enum : int { UNIT_STAGED = 0, UNIT_GATHER = 1, UNIT_END = 2 };

#ifdef DSVC_TRACE
// Debug build only (-DDSVC_TRACE): per-CTA, per-unit clock stamps of the three roles, dumped by
// the launcher to $DSVC_WARP_TRACE.  Never compiled into the product library.
constexpr int TRACE_UNITS = 24, TRACE_FIELDS = 16;
__device__ long long* g_trace = nullptr;
#define DSVC_TR(unit, field, val)                                                              \
    do {                                                                                       \
        if (g_trace && (unit) < TRACE_UNITS)                                                   \
            g_trace[((size_t)blockIdx.x * TRACE_UNITS + (unit)) * TRACE_FIELDS + (field)] = (val); \
    } while (0)
#else
#define DSVC_TR(unit, field, val) do { } while (0)
#endif

// What the scout tells the issuer and the consumers about one piece of work.
struct UnitDesc {
    int mode;                // UNIT_*
    int tx0, ty0, b;         // tile origin (pixels), batch item
    int c_begin, c_end;      // channel range
    int bx0, by0, nchunks;   // staged box origin (source pixels), 8-row chunks to load
    int rx0, rx1, ry0, ry1;  // the part of the tile this descriptor covers (tile coordinates)
    int fast;                // whole interior tile, no tap clamped at the right border
    int pad[2];
};

struct Schedule {
    int tiles_x, tiles_y;
    int full_tiles;   // units [0, full_tiles) are whole tiles, all channels
    int tail_split;   // later tiles are cut into this many channel ranges ...
    int cper;         // ... of this many channels
    int total_units;
    int strip;        // tiles per strip row (0: plain row-major tile order)
};

// Source coordinates of the pixels of rectangle [rx0,rx1) x [ry0,ry1) of the tile at (tx0, ty0),
// written to `coords` (tile-relative [y][x], TW wide) for the consumers, and their bounding box
// (north-west tap positions), warp-reduced (all lanes return the same box).  A lane owns the
// columns lane, lane + 32, ... of the rectangle (rx0, rx1 are multiples of 32) and walks down
// the rows RB at a time with all 2*RB flow loads in flight; min / max run on the float
// coordinates (floor is monotone, so floor(min) = min(floor)).
template <int TW, int RB>
__device__ __forceinline__ void scout_bbox(const float* __restrict__ fl, const float* __restrict__ lin_x,
                                           const float* __restrict__ lin_y, const WarpParams& p,
                                           int tx0, int ty0, int rx0, int rx1, int ry0, int ry1,
                                           int lane, float2* __restrict__ coords, int& mnx, int& mxx,
                                           int& mny, int& mxy) {
    const size_t plane = (size_t)p.H * p.W;
    float lox = 3.0e38f, hix = -1.0f, loy = 3.0e38f, hiy = -1.0f;  // coordinates are in [0, size-1]
    for (int xx = rx0 + lane; xx < rx1; xx += 32) {
        const int x = tx0 + xx;
        const bool xok = x < p.W;
        const int xc = min(x, p.W - 1);
        const float lx = __ldg(lin_x + xc);
        for (int r0 = ry0; r0 < ry1; r0 += RB) {
            float fx[RB], fy[RB];
            const float* fp = fl + (size_t)min(ty0 + r0, p.H - 1) * p.W + xc;
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                // rows past the rectangle / image re-read the last legal row (discarded below)
                const float* q = (r0 + j < ry1 && ty0 + r0 + j < p.H) ? fp + (size_t)j * p.W : fp;
                fx[j] = __ldg(q);
                fy[j] = __ldg(q + plane);
            }
#pragma unroll
            for (int j = 0; j < RB; ++j) {
                const int yy = r0 + j, y = ty0 + yy;
                if (xok && yy < ry1 && y < p.H) {
                    const float ix = source_coord(lx, fx[j], p.sx, p.inv_sx, p.flow_mode, p.W);
                    const float iy = source_coord(__ldg(lin_y + y), fy[j], p.sy, p.inv_sy, p.flow_mode, p.H);
                    coords[yy * TW + xx] = make_float2(ix, iy);
                    lox = fminf(lox, ix); hix = fmaxf(hix, ix);
                    loy = fminf(loy, iy); hiy = fmaxf(hiy, iy);
                }
            }
        }
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lox = fminf(lox, __shfl_xor_sync(0xffffffffu, lox, o));
        hix = fmaxf(hix, __shfl_xor_sync(0xffffffffu, hix, o));
        loy = fminf(loy, __shfl_xor_sync(0xffffffffu, loy, o));
        hiy = fmaxf(hiy, __shfl_xor_sync(0xffffffffu, hiy, o));
    }
    if (hix < 0.0f) {  // no pixel of the rectangle is inside the image
        mnx = mny = INT_MAX;
        mxx = mxy = INT_MIN;
    } else {
        mnx = (int)floorf(lox); mxx = (int)floorf(hix);
        mny = (int)floorf(loy); mxy = (int)floorf(hiy);
    }
}

template <class Cfg>
__global__ void __maxnreg__(Cfg::MAXREG)
warp_fwd_persist_kernel(const __grid_constant__ CUtensorMap tmap, const __grid_constant__ CUtensorMap tmap2,
                        const float* __restrict__ in, const float* __restrict__ in2,
                        const float* __restrict__ flow, float* __restrict__ out, float* __restrict__ out2, int C1,
                        const float* __restrict__ lin_x, const float* __restrict__ lin_y,
                        WarpParams p, Schedule sch, WarpSched* __restrict__ sched) {
    constexpr int TW = Cfg::TW, TH = Cfg::TH, BW = Cfg::BW, CC = Cfg::CC, STAGES = Cfg::STAGES;
    constexpr int RPW = Cfg::ROWS_PER_WARP, XH = Cfg::XH, PPT = Cfg::PPT, ND = Cfg::NDESC;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    float* stage_buf = reinterpret_cast<float*>(smem_raw);
    __shared__ __align__(8) uint64_t full_bar[STAGES];
    __shared__ __align__(8) uint64_t empty_bar[STAGES];
    __shared__ __align__(8) uint64_t desc_full[ND];
    __shared__ __align__(8) uint64_t desc_empty[ND];
    __shared__ __align__(16) UnitDesc desc[ND];
    __shared__ __align__(16) float2 coords[ND][TW * TH];  // per-pixel source coordinates of a unit

    pdl_prologue();  // (common.cuh: nothing of the predecessor's output is touched before this)
    // (warp index broadcast from lane 0 so that the compiler treats role branches as warp-uniform)
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
    const size_t plane = (size_t)p.H * p.W;
    // shared-window addresses of the barrier arrays (8 bytes per barrier)
    // (made opaque so that they live in registers instead of being re-derived per use)
    uint32_t full0 = tma::smem_u32(full_bar), empty0 = tma::smem_u32(empty_bar);
    uint32_t dfull0 = tma::smem_u32(desc_full), dempty0 = tma::smem_u32(desc_empty);
    uint32_t sbase0 = tma::smem_u32(stage_buf);
    asm volatile("" : "+r"(full0), "+r"(empty0), "+r"(dfull0), "+r"(dempty0), "+r"(sbase0));

    if (threadIdx.x == 0) {
#pragma unroll
        for (int s = 0; s < STAGES; ++s) {
            tma::mbar_init(&full_bar[s], 1);
            tma::mbar_init(&empty_bar[s], Cfg::CONSUMER_WARPS);
        }
#pragma unroll
        for (int d = 0; d < ND; ++d) {
            tma::mbar_init(&desc_full[d], 1);
            tma::mbar_init(&desc_empty[d], Cfg::CONSUMER_WARPS + 1);  // consumers + issuer
        }
        tma::fence_barrier_init();
    }
    __syncthreads();

    if (warp == Cfg::SCOUT_WARP) {
        // ------------------------------------------------------------------ scout
        int du = 0;
        bool have_slot = false;
        // a descriptor slot (descriptor + coordinate buffer) is written only after every reader
        // of its previous contents has released it
        auto acquire = [&]() {
            if (!have_slot && du >= ND) tma::mbar_wait(dempty0 + 8u * (du % ND), ((du / ND) - 1) & 1);
            have_slot = true;
        };
        auto post = [&](const UnitDesc& d) {
            const int slot = du % ND;
            if (lane == 0) desc[slot] = d;
            __syncwarp();  // every lane's coordinate stores precede the release below
            if (lane == 0) tma::mbar_arrive(dfull0 + 8u * slot);
            ++du;
            have_slot = false;
        };
        for (;;) {
            int u = 0;
            if (lane == 0) DSVC_TR(du, 0, clock64());
            if (lane == 0) u = atomicAdd(&sched->next, 1);
            u = __shfl_sync(0xffffffffu, u, 0);
            if (lane == 0) DSVC_TR(du, 1, clock64());
            if (u >= sch.total_units) break;
            UnitDesc d;
            int tile;
            if (u < sch.full_tiles) {
                tile = u;
                d.c_begin = 0;
                d.c_end = p.C;
            } else {
                const int v = u - sch.full_tiles;
                tile = sch.full_tiles + v / sch.tail_split;
                d.c_begin = (v % sch.tail_split) * sch.cper;
                d.c_end = min(p.C, d.c_begin + sch.cper);
                if (d.c_begin >= d.c_end) continue;
            }
            {
                // tile order: row-major inside vertical strips of `strip` tiles, so that both the
                // left/right and the above/below neighbours of a tile are in flight with it
                const int per_img = sch.tiles_x * sch.tiles_y;
                d.b = tile / per_img;
                int r = tile - d.b * per_img, tx, ty;
                if (sch.strip > 0) {
                    const int per_strip = sch.strip * sch.tiles_y;
                    const int s = r / per_strip;
                    r -= s * per_strip;
                    const int sw = min(sch.strip, sch.tiles_x - s * sch.strip);  // last strip may be narrower
                    ty = r / sw;
                    tx = s * sch.strip + (r - ty * sw);
                } else {
                    ty = r / sch.tiles_x;
                    tx = r - ty * sch.tiles_x;
                }
                d.tx0 = tx * TW;
                d.ty0 = ty * TH;
            }
            d.pad[0] = d.pad[1] = 0;
            const float* fl = flow + (size_t)d.b * 2 * plane;
            // a descriptor for rectangle [rx0,rx1) x [ry0,ry1) of the tile; false if it is the
            // whole tile and does not fit (the caller then posts the quadrants)
            auto describe = [&](int rx0, int rx1, int ry0, int ry1, bool whole) -> bool {
                int mnx, mxx, mny, mxy;
                acquire();
                scout_bbox<TW, Cfg::SCOUT_RB>(fl, lin_x, lin_y, p, d.tx0, d.ty0, rx0, rx1, ry0, ry1, lane,
                                             coords[du % ND], mnx, mxx, mny, mxy);
                if (mnx > mxx) return true;  // rectangle entirely outside the image: nothing to do
                // taps reach x0+1 / y0+1 (clamped to the image); TMA tiled loads need a 16-byte
                // aligned start along x (an unaligned coordinate raises "illegal instruction")
                const int bx0 = mnx & ~3;
                const int bw = min(mxx + 1, p.W - 1) - bx0 + 1;
                const int bh = min(mxy + 1, p.H - 1) - mny + 1;
                const bool fits = bw <= BW && bh <= Cfg::BHMAX;
                if (!fits && whole) return false;
                d.mode = fits ? UNIT_STAGED : UNIT_GATHER;
                d.bx0 = bx0;
                d.by0 = mny;
                d.nchunks = (bh + Cfg::ROWCHUNK - 1) / Cfg::ROWCHUNK;
                d.rx0 = rx0; d.rx1 = rx1; d.ry0 = ry0; d.ry1 = ry1;
                d.fast = whole && d.tx0 + TW <= p.W && d.ty0 + TH <= p.H && mxx + 1 < p.W;
                if (lane == 0) DSVC_TR(du, 2, clock64());
                if (lane == 0) DSVC_TR(du, 12, (long long)u);
                post(d);
                if (lane == 0) DSVC_TR(du - 1, 3, clock64());
                return true;
            };
            if (!describe(0, TW, 0, TH, true)) {
#pragma unroll 1
                for (int q = 0; q < Cfg::QX * Cfg::QY; ++q) {
                    const int qx = (q % Cfg::QX) * (TW / Cfg::QX), qy = (q / Cfg::QX) * (TH / Cfg::QY);
                    describe(qx, qx + TW / Cfg::QX, qy, qy + TH / Cfg::QY, false);
                }
            }
        }
        UnitDesc e{};
        e.mode = UNIT_END;
        acquire();
        post(e);
        if (lane == 0) {
            // the last scout to run dry leaves the scheduler state zeroed for the next launch
            if (atomicAdd(&sched->exited, 1) == (int)gridDim.x - 1) {
                sched->next = 0;
                __threadfence();
                sched->exited = 0;
            }
        }
        return;
    }

    if (warp == Cfg::ISSUER_WARP) {
        // ------------------------------------------------------------------ issuer
        if (lane != 0) return;
        uint32_t it = 0;
        for (int du = 0;; ++du) {
            const int slot = du % ND;
            DSVC_TR(du, 4, clock64());
            tma::mbar_wait(dfull0 + 8u * slot, (du / ND) & 1);
            const UnitDesc d = desc[slot];
            tma::mbar_arrive(dempty0 + 8u * slot);
            DSVC_TR(du, 5, clock64());
            if (d.mode == UNIT_END) break;
            if (d.mode != UNIT_STAGED) continue;
            long long tw_empty = 0;
            const int ngroups = (d.c_end - d.c_begin + CC - 1) / CC;
            const uint32_t tx_bytes = (uint32_t)d.nchunks * Cfg::CHUNK_FLOATS * sizeof(float);
            // channels [0, C1) live in the first tensor, [C1, p.C) in the second (same flow, same
            // coordinates: e.g. the frame warp of video_model.py:37 riding on the feature warp)
            for (int g = 0; g < ngroups; ++g, ++it) {
                const int cg = d.c_begin + g * CC;
                const bool second = cg >= C1;
                const CUtensorMap* tm = second ? &tmap2 : &tmap;
                const int pl = second ? d.b * (p.C - C1) + (cg - C1) : d.b * C1 + cg;
                const uint32_t s = it % STAGES;
#ifdef DSVC_TRACE
                const long long te0 = clock64();
#endif
                if (it >= STAGES) tma::mbar_wait(empty0 + 8u * s, ((it / STAGES) - 1) & 1);
#ifdef DSVC_TRACE
                tw_empty += clock64() - te0;
#endif
                tma::mbar_arrive_expect_tx(full0 + 8u * s, tx_bytes);
                const uint32_t dst = sbase0 + s * (uint32_t)(Cfg::STAGE_FLOATS * 4);
                for (int k = 0; k < d.nchunks; ++k)
                    tma::load_3d(dst + (uint32_t)k * (Cfg::CHUNK_FLOATS * 4), tm, d.bx0, pl,
                                 d.by0 + k * Cfg::ROWCHUNK, full0 + 8u * s);
            }
            DSVC_TR(du, 6, clock64());
            DSVC_TR(du, 7, tw_empty);
            (void)tw_empty;
        }
        return;
    }

    // ---------------------------------------------------------------------- consumers
    uint32_t it = 0;
    for (int du = 0;; ++du) {
        const int slot = du % ND;
        if (threadIdx.x == 0) DSVC_TR(du, 8, clock64());
        tma::mbar_wait(dfull0 + 8u * slot, (du / ND) & 1);
        const UnitDesc d = desc[slot];
        if (d.mode == UNIT_END) break;
        long long tw_full = 0;
        (void)tw_full;

        // per-pixel source coordinates: computed once by the scout, reused for every channel
        float ixs[PPT], iys[PPT];
        bool valid[PPT];
        uint32_t vmask = 0;
#pragma unroll
        for (int r = 0; r < RPW; ++r) {
#pragma unroll
            for (int h = 0; h < XH; ++h) {
                const int k = r * XH + h;
                const int xx = h * 32 + lane, yy = warp * RPW + r;
                valid[k] = xx >= d.rx0 && xx < d.rx1 && yy >= d.ry0 && yy < d.ry1 &&
                           d.tx0 + xx < p.W && d.ty0 + yy < p.H;
                vmask |= valid[k] ? 1u << k : 0u;
                const float2 c = valid[k] ? coords[slot][yy * TW + xx] : make_float2(0.0f, 0.0f);
                ixs[k] = c.x;
                iys[k] = c.y;
            }
        }
        __syncwarp();
        if (lane == 0) tma::mbar_arrive(dempty0 + 8u * slot);  // descriptor and coordinates are in registers
        if (threadIdx.x == 0) DSVC_TR(du, 9, clock64());

        if (d.mode == UNIT_GATHER) {
            // rectangle whose taps do not fit the staging box: read-only-path gather
#pragma unroll 1
            for (int k = 0; k < PPT; ++k) {
                if (!((vmask >> k) & 1u)) continue;
                const int x = d.tx0 + (k % XH) * 32 + lane, y = d.ty0 + warp * RPW + k / XH;
                if (d.c_begin < C1) {
                    WarpParams pa = p;
                    pa.C = C1;
                    gather_pixel<8>(in, flow, out, lin_x, lin_y, pa, d.b, x, y, d.c_begin, min(d.c_end, C1));
                }
                if (d.c_end > C1) {
                    WarpParams pb = p;
                    pb.C = p.C - C1;
                    gather_pixel<8>(in2, flow, out2, lin_x, lin_y, pb, d.b, x, y, max(d.c_begin, C1) - C1, d.c_end - C1);
                }
            }
            continue;
        }

        // gather from the staged box, store coalesced rows.  Shared-memory element (row ry,
        // channel c, column rx) of a stage lives at
        //   (ry >> 3) * CHUNK_FLOATS + (ry & 7) * ROW_PITCH + c * BW + rx.
        const int ngroups = (d.c_end - d.c_begin + CC - 1) / CC;
        const int nch = d.c_end - d.c_begin;
        const size_t toff = (size_t)(d.ty0 + warp * RPW) * p.W + d.tx0 + lane;
        auto out_plane = [&](int c) -> float* {  // channel c of batch item d.b, at the thread's tile offset
            return (c < C1 ? out + (size_t)(d.b * C1 + c) * plane
                           : out2 + (size_t)(d.b * (p.C - C1) + (c - C1)) * plane) + toff;
        };
        float* obase = out_plane(d.c_begin);
        if (d.fast) {
            // interior tile, every east tap inside the image: byte addresses, immediate offsets
            float wnw[PPT], wne[PPT], wsw[PPT], wse[PPT];
            uint32_t a_n[PPT], a_s[PPT];
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                const Taps t = make_taps(ixs[k], iys[k], p.W, p.H);
                const int rx = t.x0 - d.bx0, ry = t.y0 - d.by0;
                const int ry1 = ry + (t.y1ok ? 1 : 0);
                a_n[k] = 4u * (uint32_t)((ry >> 3) * Cfg::CHUNK_FLOATS + (ry & 7) * Cfg::ROW_PITCH + rx);
                a_s[k] = 4u * (uint32_t)((ry1 >> 3) * Cfg::CHUNK_FLOATS + (ry1 & 7) * Cfg::ROW_PITCH + rx);
                wnw[k] = t.nw;
                wne[k] = t.ne;
                wsw[k] = t.y1ok ? t.sw : 0.0f;
                wse[k] = t.y1ok ? t.se : 0.0f;
            }
            // outputs are written once and never read here: first in line for L2 eviction, so that
            // they do not push out the input rows neighbouring tiles are about to re-read (-5 %)
            const uint64_t st_pol = tma::policy_evict_first();
            // output row pointers of the current channel; one plane further per channel
            float* orow[RPW];
#pragma unroll
            for (int r = 0; r < RPW; ++r) orow[r] = obase + (size_t)r * p.W;
            // channels whose 4*PPT taps are loaded back to back before any arithmetic
            constexpr int CB = (CC * PPT * 4 <= 32) ? CC : 1;
            if (threadIdx.x == 0) DSVC_TR(du, 10, clock64());
            for (int g = 0; g < ngroups; ++g, ++it) {
                const uint32_t s = it % STAGES;
#ifdef DSVC_TRACE
                const long long tf0 = clock64();
#endif
                tma::mbar_wait(full0 + 8u * s, (it / STAGES) & 1);
#ifdef DSVC_TRACE
                tw_full += clock64() - tf0;
#endif
                const uint32_t sbase = sbase0 + s * (uint32_t)(Cfg::STAGE_FLOATS * 4);
                if (g > 0 && d.c_begin + g * CC == C1) {  // the unit crosses into the second tensor
                    float* ob2 = out_plane(C1);
#pragma unroll
                    for (int r = 0; r < RPW; ++r) orow[r] = ob2 + (size_t)r * p.W;
                }
                uint32_t tn[PPT], ts[PPT];
#pragma unroll
                for (int k = 0; k < PPT; ++k) { tn[k] = a_n[k] + sbase; ts[k] = a_s[k] + sbase; }
                tma::static_for<CC / CB>([&](auto bb) {
                    constexpr int c0 = decltype(bb)::value * CB;
                    float v[CB][PPT][4];
                    tma::static_for<CB>([&](auto jj) {
                        constexpr int j = decltype(jj)::value, c = c0 + j;
#pragma unroll
                        for (int k = 0; k < PPT; ++k) {
                            v[j][k][0] = tma::lds_imm<c * BW * 4>(tn[k]);
                            v[j][k][1] = tma::lds_imm<c * BW * 4 + 4>(tn[k]);
                            v[j][k][2] = tma::lds_imm<c * BW * 4>(ts[k]);
                            v[j][k][3] = tma::lds_imm<c * BW * 4 + 4>(ts[k]);
                        }
                    });
                    tma::static_for<CB>([&](auto jj) {
                        constexpr int j = decltype(jj)::value;
                        if (g * CC + c0 + j < nch) {  // warp-uniform (odd channel counts)
#pragma unroll
                            for (int r = 0; r < RPW; ++r) {
                                tma::static_for<XH>([&](auto hh) {
                                    constexpr int h = decltype(hh)::value;
                                    const int k = r * XH + h;
                                    float acc = __fmul_rn(v[j][k][0], wnw[k]);
                                    acc = fmaf(v[j][k][1], wne[k], acc);
                                    acc = fmaf(v[j][k][2], wsw[k], acc);
                                    acc = fmaf(v[j][k][3], wse[k], acc);
                                    st_hint_imm<h * 128>(orow[r], acc, st_pol);
                                });
                                orow[r] += plane;
                            }
                        }
                    });
                });
                __syncwarp();
                if (lane == 0) tma::mbar_arrive(empty0 + 8u * s);
            }
            if (threadIdx.x == 0) { DSVC_TR(du, 11, clock64()); DSVC_TR(du, 13, tw_full); DSVC_TR(du, 14, (long long)ngroups); }
        } else {
            // edge tile, quadrant, or taps clamped at the right image border: generic loop
            float wnw[PPT], wne[PPT], wsw[PPT], wse[PPT];
            int off_n[PPT], off_s[PPT], dxs[PPT];
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                const Taps t = make_taps(ixs[k], iys[k], p.W, p.H);
                const int rx = t.x0 - d.bx0, ry = t.y0 - d.by0;
                const int ry1 = ry + (t.y1ok ? 1 : 0);
                off_n[k] = (ry >> 3) * Cfg::CHUNK_FLOATS + (ry & 7) * Cfg::ROW_PITCH + rx;
                off_s[k] = (ry1 >> 3) * Cfg::CHUNK_FLOATS + (ry1 & 7) * Cfg::ROW_PITCH + rx;
                dxs[k] = t.x1ok ? 1 : 0;
                // a tap outside the image contributes nothing (ATen skips it): zero its weight,
                // its (clamped) address stays inside the staged box
                wnw[k] = t.nw;
                wne[k] = t.x1ok ? t.ne : 0.0f;
                wsw[k] = t.y1ok ? t.sw : 0.0f;
                wse[k] = (t.x1ok && t.y1ok) ? t.se : 0.0f;
                if (!valid[k]) { off_n[k] = off_s[k] = 0; dxs[k] = 0; }
            }
            for (int g = 0; g < ngroups; ++g, ++it) {
                const uint32_t s = it % STAGES;
                tma::mbar_wait(full0 + 8u * s, (it / STAGES) & 1);
                const float* sb = stage_buf + (size_t)s * Cfg::STAGE_FLOATS;
#pragma unroll
                for (int c = 0; c < CC; ++c) {
                    const int ch = g * CC + c;
                    if (ch < nch) {
                        const float* sc = sb + c * BW;
                        float* oc = out_plane(d.c_begin + ch);
#pragma unroll
                        for (int r = 0; r < RPW; ++r)
#pragma unroll
                            for (int h = 0; h < XH; ++h) {
                                const int k = r * XH + h;
                                const float a = sc[off_n[k]], bq = sc[off_n[k] + dxs[k]];
                                const float cq = sc[off_s[k]], dq = sc[off_s[k] + dxs[k]];
                                float acc = __fmul_rn(a, wnw[k]);
                                acc = fmaf(bq, wne[k], acc);
                                acc = fmaf(cq, wsw[k], acc);
                                acc = fmaf(dq, wse[k], acc);
                                if (valid[k]) st_stream1(oc + (size_t)r * p.W + h * 32, acc);
                            }
                    }
                }
                __syncwarp();
                if (lane == 0) tma::mbar_arrive(empty0 + 8u * s);
            }
        }
    }
}

}  // namespace dsvc

using namespace dsvc;

// ------------------------------------------------------------------------ host side
template <class Cfg>
static int launch_persist(const float* input, const float* flow, float* out, const float* lin_x,
                          const float* lin_y, const WarpParams& p, void* workspace,
                          size_t workspace_bytes, cudaStream_t st, const float* input2 = nullptr,
                          float* out2 = nullptr, int C2 = 0) {
    // p.C counts the channels of both tensors; the first one has C1 = p.C - C2 of them
    const int C1 = p.C - C2;
    if (C2 > 0 && (C1 % Cfg::CC != 0 || !input2 || !out2)) return -1;
    auto encode = tensor_map_encoder();
    if (!encode) return -1;
    if (!workspace || workspace_bytes < sizeof(WarpSched) || !aligned16(workspace))
        return -1;  // no scheduler state: the caller uses the gather kernel
    CUtensorMap tm;
    // tensor viewed as (x, plane, y): the box lands in shared memory as [8 rows][CC][BW]
    const cuuint64_t gdim[3] = {(cuuint64_t)p.W, (cuuint64_t)p.B * C1, (cuuint64_t)p.H};
    const cuuint64_t gstride[2] = {(cuuint64_t)p.H * p.W * 4, (cuuint64_t)p.W * 4};
    const cuuint32_t box[3] = {(cuuint32_t)Cfg::BW, (cuuint32_t)Cfg::CC, (cuuint32_t)Cfg::ROWCHUNK};
    const cuuint32_t estr[3] = {1, 1, 1};
    const CUtensorMapL2promotion promo = CU_TENSOR_MAP_L2_PROMOTION_L2_128B;  // (no measurable effect)
    const CUresult r = encode(&tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(input),
                              gdim, gstride, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_NONE, promo, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) return -1;
    CUtensorMap tm2 = tm;
    if (C2 > 0) {
        const cuuint64_t gdim2[3] = {(cuuint64_t)p.W, (cuuint64_t)p.B * C2, (cuuint64_t)p.H};
        if (encode(&tm2, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(input2), gdim2, gstride, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, promo,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
            return -1;
    }
    // per-device set-up (function attributes and SM counts belong to the device, not the process)
    static unsigned long long attr_set = 0;
    static int sms_of[64];
    // schedule parameters (measured on B200): the last wave's tiles are cut into 2
    // channel ranges, tiles are ordered row-major inside 20-tile strips.  -DDSVC_TUNE builds read
    // $DSVC_WARP_TAIL_SPLIT / _TAIL_PCT / _STRIP once per device for re-tuning; the product
    // library has no environment knobs in its launch path.
    static int env_split = 0, env_tail_pct = 100, env_strip = 20;
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
    if (!((attr_set >> dev) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(warp_fwd_persist_kernel<Cfg>,
                                             cudaFuncAttributeMaxDynamicSharedMemorySize,
                                             (int)Cfg::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
            n = DSVC_NUM_SMS;
        sms_of[dev] = n;
#ifdef DSVC_TUNE
        if (const char* e2 = getenv("DSVC_WARP_TAIL_SPLIT")) env_split = atoi(e2);
        if (const char* e2 = getenv("DSVC_WARP_TAIL_PCT")) env_tail_pct = atoi(e2);
        if (const char* e2 = getenv("DSVC_WARP_STRIP")) env_strip = atoi(e2);
#endif
        attr_set |= 1ull << dev;
    }
    const int num_sms = sms_of[dev];
    const int slots = Cfg::MINB * num_sms;  // __launch_bounds__(THREADS, MINB)
    Schedule sch;
    sch.tiles_x = (p.W + Cfg::TW - 1) / Cfg::TW;
    sch.tiles_y = (p.H + Cfg::TH - 1) / Cfg::TH;
    const long long ntiles = (long long)sch.tiles_x * sch.tiles_y * p.B;
    if (ntiles > (1ll << 28)) return -1;
    // the last `tail` tiles (about one wave) are cut into channel ranges of >= 8 channels
    int split = env_split > 0 ? env_split : 2;
    while (split > 1 && p.C / split < 8) split >>= 1;
    const long long tail = std::min<long long>(ntiles, (long long)slots * env_tail_pct / 100);
    sch.tail_split = split;
    sch.cper = ((p.C + split - 1) / split + Cfg::CC - 1) / Cfg::CC * Cfg::CC;
    // 20-tile (1280-pixel) strips: at 3840 wide the plain row-major order keeps a tile's vertical
    // neighbours 60 units apart and costs 19 % (924 -> 776 us at 2176x3840 C=64; strips of 10..30
    // tiles measure the same, 1080p is unchanged)
    sch.strip = env_strip;
    sch.full_tiles = (int)(ntiles - tail);
    sch.total_units = (int)(sch.full_tiles + tail * split);
    const int grid = (int)std::min<long long>(slots, sch.total_units);
#ifdef DSVC_TRACE
    const char* trace_path = getenv("DSVC_WARP_TRACE");
    long long* trace = nullptr;
    const size_t trace_n = (size_t)grid * TRACE_UNITS * TRACE_FIELDS;
    if (trace_path) {
        cudaMalloc(&trace, trace_n * sizeof(long long));
        cudaMemset(trace, 0, trace_n * sizeof(long long));
        cudaMemcpyToSymbol(g_trace, &trace, sizeof(trace));
    }
#endif
    launch_pdl(warp_fwd_persist_kernel<Cfg>, dim3(grid), dim3(Cfg::THREADS), Cfg::SMEM_BYTES, st, tm, tm2, input, input2,
               flow, out, out2, C1, lin_x, lin_y, p, sch, static_cast<WarpSched*>(workspace));
#ifdef DSVC_TRACE
    if (trace) {
        cudaStreamSynchronize(st);
        std::vector<long long> h(trace_n);
        cudaMemcpy(h.data(), trace, trace_n * sizeof(long long), cudaMemcpyDeviceToHost);
        if (FILE* f = fopen(trace_path, "wb")) {
            const int hdr[4] = {grid, TRACE_UNITS, TRACE_FIELDS, 0};
            fwrite(hdr, sizeof(int), 4, f);
            fwrite(h.data(), sizeof(long long), trace_n, f);
            fclose(f);
        }
        long long* null = nullptr;
        cudaMemcpyToSymbol(g_trace, &null, sizeof(null));
        cudaFree(trace);
    }
#endif
    return (int)cudaGetLastError();
}

// returns -1 when the shape is not eligible (caller uses the gather kernel)
int dsvc_warp_fwd_persist_launch(const float* input, const float* flow, float* out,
                                 const float* lin_x, const float* lin_y, const WarpParams& p,
                                 bool force, void* workspace, size_t workspace_bytes,
                                 cudaStream_t st, const float* input2, float* out2, int C2) {
    // TMA needs 16-byte aligned rows and base; small / few-channel warps gain nothing
    if (p.W % 4 != 0 || !aligned16(input) || (C2 > 0 && !aligned16(input2))) return -1;
    if (C2 > 0)  // two tensors, one flow: the default configuration only
        return launch_persist<PersistCfg<64, 16, 96, 32, 1, 6, 2, 72>>(input, flow, out, lin_x, lin_y, p, workspace,
                                                                      workspace_bytes, st, input2, out2, C2);
    if (!force && (p.C < 8 || p.W < 64 || p.H < 32)) return -1;
    if ((long long)p.B * p.C > (1ll << 30)) return -1;
#ifdef DSVC_TUNE
    // tile / box / stage / register-cap variants for re-tuning (DESIGN.md 4.2 lists what was measured);
    // never compiled into the product library
    static int cfg = -1;
    if (cfg < 0) {
        const char* e = getenv("DSVC_TMA_CFG");  // tuning knob
        cfg = e ? atoi(e) : 0;
    }
    switch (cfg) {
        case 1: return launch_persist<PersistCfg<64, 32, 80, 48, 2, 2>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 2: return launch_persist<PersistCfg<64, 32, 80, 48, 4, 2>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 3: return launch_persist<PersistCfg<64, 32, 96, 48, 2, 3>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 4: return launch_persist<PersistCfg<64, 16, 80, 32, 2, 4>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 5: return launch_persist<PersistCfg<64, 16, 80, 32, 2, 5>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 6: return launch_persist<PersistCfg<32, 32, 48, 48, 2, 5>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 7: return launch_persist<PersistCfg<64, 16, 80, 32, 2, 3>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 9: return launch_persist<PersistCfg<64, 16, 80, 32, 2, 3, 3>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 10: return launch_persist<PersistCfg<64, 32, 80, 48, 2, 3>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 11: return launch_persist<PersistCfg<128, 16, 160, 32, 1, 4>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 12: return launch_persist<PersistCfg<128, 8, 160, 24, 1, 5>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 13: return launch_persist<PersistCfg<64, 16, 96, 32, 1, 6>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 14: return launch_persist<PersistCfg<128, 16, 160, 32, 1, 5>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 15: return launch_persist<PersistCfg<64, 16, 80, 32, 1, 6>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 16: return launch_persist<PersistCfg<64, 16, 96, 32, 1, 7>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 17: return launch_persist<PersistCfg<64, 16, 96, 32, 1, 4>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 18: return launch_persist<PersistCfg<64, 16, 80, 32, 1, 8>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 19: return launch_persist<PersistCfg<64, 16, 80, 32, 1, 5>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 20: return launch_persist<PersistCfg<64, 16, 80, 32, 1, 7>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 8: return launch_persist<PersistCfg<64, 16, 80, 32, 2, 2>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 22: return launch_persist<PersistCfg<64, 16, 96, 32, 1, 6, 2, 72>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 23: return launch_persist<PersistCfg<64, 16, 96, 32, 1, 6, 2, 80>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 21: return launch_persist<PersistCfg<64, 16, 80, 32, 2, 4>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        case 24: return launch_persist<PersistCfg<64, 16, 96, 32, 1, 6>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
        default: return launch_persist<PersistCfg<64, 16, 96, 32, 1, 6, 2, 72>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
    }
#else
    // 64 x 16 tiles, 96 x 32 staging box, one channel plane per stage, 6 stages, 2 CTAs per SM, 72 registers
    return launch_persist<PersistCfg<64, 16, 96, 32, 1, 6, 2, 72>>(input, flow, out, lin_x, lin_y, p, workspace, workspace_bytes, st);
#endif
}
