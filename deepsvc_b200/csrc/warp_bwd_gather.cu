// Backward of the bilinear warp as a destination-owned GATHER (NCHW fp32, sm_100a).
//
// Replaces the autograd of /root/reference/modules.py:25-62 (ATen grid_sampler_2d_backward via
// Learner.py:1343) for the bandwidth-critical call: the 64-ch feature warp of modules.py:429
// needs grad_input and grad_flow in training.
//
// grad_input is the TRANSPOSED warp applied to grad_out.  Instead of scattering every output
// pixel's four taps (global atomics: 249 M L2 reduction sectors per 1080p launch; or the r01
// shared-memory CSR scatter with its out-box, 3-way bank conflicts and zero-filled
// grad_input), a CTA OWNS a 64 x 16 tile of grad_input and gathers into it:
//   * once per tile (the geometry does not depend on the channel) it scans a 96 x 48 search
//     region of output pixels -- the tile moved by minus the flow at its centre, +-16 pixels --
//     and records every tap that lands in the tile as a (weight, source pixel) pair in the
//     destination element's list: the first 8 pairs of an element in an ELL table that the
//     element's owner thread then keeps in REGISTERS, later ones (compressive flows, border
//     pile-ups: ~3 % of the pairs) in a shared-memory list sorted by owner;
//   * per channel the bounding box of the contributing pixels is TMA-staged from grad_out
//     (<= 96 x 40) and every thread sums its four elements' lists -- lane = destination column,
//     so a warp's gathers walk consecutive columns of the box -- and writes grad_input with
//     plain coalesced stores: no atomics, no zero-fill of grad_input, one store per element;
//   * grad_flow: as in the forward kernel the four input taps of the thread's own output
//     pixels come from the TMA-staged source box of the input; grad_out of the own pixel is
//     read from global memory one channel ahead; the channel sum lives in registers.
// Taps the scan cannot see (an output pixel outside its destination tile's search region) and
// tiles whose boxes or lists do not fit (wild flows) are completed by the FIX-UP launch that
// follows in the stream: one thread per output pixel recomputes the same predicate (plus a
// per-tile flag byte the main kernel wrote) and adds exactly the missing taps with global
// atomics.  For SpyNet-like flows the fix-up launch finds ~7e-6 of the taps.
#include <algorithm>
#include <climits>
#include <cstdlib>

#include "tma_utils.cuh"
#include "warp_bwd_common.cuh"

namespace dsvc {

namespace bg {
constexpr int TW = 64, TH = 16;               // destination tile = the CTA's output tile
constexpr int THREADS = 256;
constexpr int PPT = TW * TH / THREADS;        // elements (and output pixels) per thread: 4
constexpr int NE = TW * TH;
constexpr int SMX = 16, SMY = 16;             // search margins around the moved tile
constexpr int SRW = TW + 2 * SMX, SRH = TH + 2 * SMY;  // 96 x 48 search region
constexpr int SPT = SRW * SRH / THREADS;      // search pixels per thread: 18
constexpr int GBW = 96, GBH = 40;             // staged grad_out box (floats x rows)
constexpr int IBW = 96, IBH = 32;             // staged input box, as the forward kernel's
constexpr int ROWCHUNK = 8;                   // rows per TMA box
constexpr int K = 8;                          // register slots per destination element
constexpr int OVF = 2048;                     // later pairs of a tile (shared-memory list)
constexpr int NS = 3;                         // load stages
constexpr int STAGE_G = GBW * GBH, STAGE_IN = IBW * IBH;
constexpr int G_BYTES = STAGE_G * 4 + 128, IN_BYTES = STAGE_IN * 4;  // grad_out box + a zero word (unused list slots)
constexpr int ZERO_OFF = STAGE_G * 4;         // the zero word of a stage, relative to its grad_out box
constexpr int IN_OFF = NS * G_BYTES;          // dynamic smem: [NS grad_out boxes][NS input boxes][list]
constexpr int OVF_OFF = IN_OFF + NS * IN_BYTES;
constexpr size_t SMEM_BYTES = (size_t)OVF_OFF + (size_t)OVF * 8;
// build-time scratch, aliased onto the (not yet loaded) stages
constexpr int ELLW_OFF = 0, ELLQ_OFF = ELLW_OFF + K * NE * 4, TMP_OFF = ELLQ_OFF + K * NE * 2,
              OBASE_OFF = TMP_OFF + OVF * 8, CNT2_OFF = OBASE_OFF + NE * 4, SCRATCH_END = CNT2_OFF + NE * 4;
static_assert(SRW * SRH % THREADS == 0, "search pixels per thread");
static_assert(SCRATCH_END <= OVF_OFF, "the build scratch aliases the load stages");
static_assert(G_BYTES % 128 == 0 && IN_BYTES % 128 == 0 && (ROWCHUNK * GBW * 4) % 128 == 0, "TMA alignment");
static_assert(SRW * SRH <= 8192 && NE <= 1024, "pair packing: 13 bits of search pixel, 10 of element");

__device__ __forceinline__ float lds(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1];" : "=f"(v) : "r"(a));
    return v;
}
template <int IMM>
__device__ __forceinline__ float lds_i(uint32_t a) {
    float v;
    asm volatile("ld.shared.f32 %0, [%1+%2];" : "=f"(v) : "r"(a), "n"(IMM));
    return v;
}
__device__ __forceinline__ float ldg_stream(const float* p) {
    float v;
    asm volatile("ld.global.nc.L1::no_allocate.f32 %0, [%1];" : "=f"(v) : "l"(p));
    return v;
}
__device__ __forceinline__ void fence_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// Origin of the search region of the destination tile at (tx0, ty0): the tile moved by minus
// the flow at its centre (x in steps of 4), widened by the margins.  The main kernel scans
// exactly these output pixels; the fix-up kernel recomputes the same origin to find the taps
// the scan could not see.
__device__ __forceinline__ void search_origin(const float* __restrict__ fl, size_t plane, const WarpParams& p,
                                              int tx0, int ty0, int& sx0, int& sy0) {
    const int cx = min(tx0 + TW / 2, p.W - 1), cy = min(ty0 + TH / 2, p.H - 1);
    float fx = __ldg(fl + (size_t)cy * p.W + cx), fy = __ldg(fl + plane + (size_t)cy * p.W + cx);
    // (NaN: fmaxf / fminf return the other operand)
    fx = fminf(fmaxf(fx, -(float)p.W), (float)p.W);
    fy = fminf(fmaxf(fy, -(float)p.H), (float)p.H);
    sx0 = tx0 - 4 * __float2int_rn(fx * 0.25f) - SMX;
    sy0 = ty0 - __float2int_rn(fy) - SMY;
}
}  // namespace bg

// grid.x = tiles * csplit; unit u -> tile u / csplit, channel range (u % csplit) * cper ...
template <bool NEED_GFLOW>
__global__ void __launch_bounds__(bg::THREADS, 2)
warp_bwd_gather_kernel(const __grid_constant__ CUtensorMap tm_in, const __grid_constant__ CUtensorMap tm_gout,
                       const float* __restrict__ gout, const float* __restrict__ in,
                       const float* __restrict__ flow, float* __restrict__ gin, float* __restrict__ gflow,
                       const float* __restrict__ lin_x, const float* __restrict__ lin_y, WarpParams p,
                       int tiles_x, int tiles_y, int csplit, int cper, unsigned char* __restrict__ flags) {
    using namespace bg;
    extern __shared__ __align__(128) unsigned char smem_raw[];
    __shared__ __align__(8) uint64_t full_bar[NS];
    __shared__ __align__(16) int cnt[NE];  // taps landing on every element of the tile
    __shared__ int red_i[8][8];
    __shared__ int scan_w[8];
    __shared__ int ovf_n;

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int unit = blockIdx.x;
    const int tile = unit / csplit, cs = unit - tile * csplit;
    const int c_begin = cs * cper, c_end = min(p.C, c_begin + cper);
    if (c_begin >= c_end) return;
    const int tx0 = (tile % tiles_x) * TW, ty0 = ((tile / tiles_x) % tiles_y) * TH;
    const int b = tile / (tiles_x * tiles_y);
    const size_t plane = (size_t)p.H * p.W;
    const bool acc_gflow = csplit > 1;
    const float* fl = flow + (size_t)b * 2 * plane;

    float* ellw = reinterpret_cast<float*>(smem_raw + ELLW_OFF);              // [K][NE] weights
    unsigned short* ellq = reinterpret_cast<unsigned short*>(smem_raw + ELLQ_OFF);  // [K][NE] search pixel
    uint2* tmp = reinterpret_cast<uint2*>(smem_raw + TMP_OFF);                // unsorted later pairs
    int* obase = reinterpret_cast<int*>(smem_raw + OBASE_OFF);                // [NE] first later pair
    int* cnt2 = reinterpret_cast<int*>(smem_raw + CNT2_OFF);                  // [NE] placement cursor
    uint2* ovf = reinterpret_cast<uint2*>(smem_raw + OVF_OFF);                // sorted later pairs (live in the loop)

    if (tid == 0) {
#pragma unroll
        for (int s = 0; s < NS; ++s) tma::mbar_init(&full_bar[s], 1);
        tma::fence_barrier_init();
        ovf_n = 0;
    }
    for (int i = tid; i < NE; i += THREADS) { cnt[i] = 0; cnt2[i] = 0; }

    // ---- own output pixels (grad_flow): k = r * 2 + h -> (xx, yy) = (h * 32 + lane, warp * 2 + r)
    BwdCoord bc[PPT];
    bool valid[PPT];
    int mnx = INT_MAX, mxx = INT_MIN, mny = INT_MAX, mxy = INT_MIN;
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        const int x = tx0 + (k & 1) * 32 + lane, y = ty0 + warp * 2 + (k >> 1);
        valid[k] = x < p.W && y < p.H;
        if (NEED_GFLOW) {
            const int xc = min(x, p.W - 1), yc = min(y, p.H - 1);
            const size_t pix = (size_t)yc * p.W + xc;
            bc[k] = bwd_coord(__ldg(lin_x + xc), __ldg(lin_y + yc), __ldg(fl + pix), __ldg(fl + plane + pix), p);
            if (valid[k]) {
                mnx = min(mnx, bc[k].t.x0); mxx = max(mxx, bc[k].t.x0);
                mny = min(mny, bc[k].t.y0); mxy = max(mxy, bc[k].t.y0);
            }
        }
    }
    __syncthreads();  // counters zeroed

    // ---- scan of the search region: every tap that lands in the tile joins its element's list
    int sx0, sy0;
    search_origin(fl, plane, p, tx0, ty0, sx0, sy0);
    int gmnx = INT_MAX, gmxx = INT_MIN, gmny = INT_MAX, gmxy = INT_MIN;  // contributing pixels (search coordinates)
#pragma unroll 3
    for (int k = 0; k < SPT; ++k) {
        const int i = tid + k * THREADS;
        const int ry = i / SRW, rx = i - ry * SRW;
        const int x = sx0 + rx, y = sy0 + ry;
        if (x < 0 || x >= p.W || y < 0 || y >= p.H) continue;
        const size_t pix = (size_t)y * p.W + x;
        const BwdCoord c = bwd_coord(__ldg(lin_x + x), __ldg(lin_y + y), __ldg(fl + pix), __ldg(fl + plane + pix), p);
        const int ex = c.t.x0 - tx0, ey = c.t.y0 - ty0;  // north-west tap, tile coordinates
        bool any = false;
        auto tap = [&](int dx, int dy, bool ok, float w) {
            const int xx = ex + dx, yy = ey + dy;
            if (!ok || (unsigned)xx >= (unsigned)TW || (unsigned)yy >= (unsigned)TH) return;
            const int e = yy * TW + xx;
            const int r = atomicAdd(&cnt[e], 1);
            if (r < K) {
                ellw[r * NE + e] = w;
                ellq[r * NE + e] = (unsigned short)i;
            } else {
                const int pos = atomicAdd(&ovf_n, 1);
                if (pos < OVF) tmp[pos] = make_uint2(__float_as_uint(w), ((uint32_t)e << 13) | (uint32_t)i);
            }
            any = true;
        };
        tap(0, 0, true, c.t.nw);
        tap(1, 0, c.t.x1ok, c.t.ne);
        tap(0, 1, c.t.y1ok, c.t.sw);
        tap(1, 1, c.t.x1ok && c.t.y1ok, c.t.se);
        if (any) {
            gmnx = min(gmnx, rx); gmxx = max(gmxx, rx);
            gmny = min(gmny, ry); gmxy = max(gmxy, ry);
        }
    }
    mnx = __reduce_min_sync(0xffffffffu, mnx); mxx = __reduce_max_sync(0xffffffffu, mxx);
    mny = __reduce_min_sync(0xffffffffu, mny); mxy = __reduce_max_sync(0xffffffffu, mxy);
    gmnx = __reduce_min_sync(0xffffffffu, gmnx); gmxx = __reduce_max_sync(0xffffffffu, gmxx);
    gmny = __reduce_min_sync(0xffffffffu, gmny); gmxy = __reduce_max_sync(0xffffffffu, gmxy);
    if (lane == 0) {
        red_i[warp][0] = mnx; red_i[warp][1] = mxx; red_i[warp][2] = mny; red_i[warp][3] = mxy;
        red_i[warp][4] = gmnx; red_i[warp][5] = gmxx; red_i[warp][6] = gmny; red_i[warp][7] = gmxy;
    }
    __syncthreads();  // lists complete
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        mnx = min(mnx, red_i[w][0]); mxx = max(mxx, red_i[w][1]);
        mny = min(mny, red_i[w][2]); mxy = max(mxy, red_i[w][3]);
        gmnx = min(gmnx, red_i[w][4]); gmxx = max(gmxx, red_i[w][5]);
        gmny = min(gmny, red_i[w][6]); gmxy = max(gmxy, red_i[w][7]);
    }
    // input box (grad_flow): as the forward kernel; TMA boxes start 16-byte aligned along x
    const int bx0 = NEED_GFLOW ? (mnx & ~3) : 0, by0 = NEED_GFLOW ? mny : 0;
    const int bw = min(mxx + 1, p.W - 1) - bx0 + 1, bh = min(mxy + 1, p.H - 1) - by0 + 1;
    const bool in_ok = !NEED_GFLOW || (mnx <= mxx && bw <= IBW && bh <= IBH);
    // grad_out box: bounding box of the contributing pixels (empty when nothing lands in the tile)
    const bool g_any = gmnx <= gmxx;
    const int gbx0 = g_any ? ((sx0 + gmnx) & ~3) : 0, gby0 = g_any ? sy0 + gmny : 0;
    const int gbw = g_any ? sx0 + gmxx - gbx0 + 1 : 0, gbh = g_any ? gmxy - gmny + 1 : 0;
    const int n_ovf = ovf_n;
    const uint32_t gsm0 = tma::smem_u32(smem_raw);  // grad_out box of stage 0; 16-bit addresses below
    const bool fast = in_ok && gbw <= GBW && gbh <= GBH && n_ovf <= OVF && gsm0 + (uint32_t)IN_OFF <= 65536u;
    if (tid == 0) flags[tile] = fast ? 0 : 1;  // (every channel range of a tile computes the same value)

    const int nch = c_end - c_begin;
    const int plane0 = b * p.C + c_begin;
    if (!fast) {
        // wild flow: this tile's grad_input is left to the fix-up launch (zeros here), grad_flow
        // is computed per pixel from global memory
        float gx[PPT], gy[PPT];
#pragma unroll
        for (int k = 0; k < PPT; ++k) gx[k] = gy[k] = 0.0f;
        for (int c = 0; c < nch; ++c) {
            const size_t cb = (size_t)(plane0 + c) * plane;
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                if (!valid[k]) continue;
                const size_t pix = (size_t)(ty0 + warp * 2 + (k >> 1)) * p.W + tx0 + (k & 1) * 32 + lane;
                gin[cb + pix] = 0.0f;
                if (NEED_GFLOW) {
                    const Taps& t = bc[k].t;
                    const float g = __ldg(gout + cb + pix);
                    const float* ip = in + cb + (size_t)t.y0 * p.W + t.x0;
                    const int dx = t.x1ok ? 1 : 0, dy = t.y1ok ? p.W : 0;
                    const float v_nw = __ldg(ip), v_ne = __ldg(ip + dx), v_sw = __ldg(ip + dy), v_se = __ldg(ip + dy + dx);
                    const float tx = fmaf(bc[k].wy1, v_se - v_sw, bc[k].wy0 * (v_ne - v_nw));
                    const float ty = fmaf(bc[k].wx1, v_se - v_ne, bc[k].wx0 * (v_sw - v_nw));
                    gx[k] = fmaf(tx, g, gx[k]);
                    gy[k] = fmaf(ty, g, gy[k]);
                }
            }
        }
        if (NEED_GFLOW) {
#pragma unroll
            for (int k = 0; k < PPT; ++k)
                if (valid[k])
                    store_gflow(gflow, p, b, (size_t)(ty0 + warp * 2 + (k >> 1)) * p.W + tx0 + (k & 1) * 32 + lane,
                                bc[k], gx[k], gy[k], acc_gflow);
        }
        return;
    }

    // ---- later pairs: exclusive scan of max(cnt - K, 0) in element order -> obase
    {
        const int4 c4 = reinterpret_cast<const int4*>(cnt)[tid];
        const int o0 = max(c4.x - K, 0), o1 = max(c4.y - K, 0), o2 = max(c4.z - K, 0), o3 = max(c4.w - K, 0);
        const int sum = o0 + o1 + o2 + o3;
        int incl = sum;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int v = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += v;
        }
        if (lane == 31) scan_w[warp] = incl;
        __syncthreads();
        int base = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) base += w < warp ? scan_w[w] : 0;
        const int s0 = base + incl - sum;
        reinterpret_cast<int4*>(obase)[tid] = make_int4(s0, s0 + o0, s0 + o0 + o1, s0 + o0 + o1 + o2);
    }
    __syncthreads();

    // ---- the thread's four elements: lists into registers (lane = destination column)
    // a pair's address: the pixel's position in the staged grad_out box of stage 0 (16 bits)
    auto box_addr = [&](uint32_t q) -> uint32_t {
        const int ry = (int)q / SRW, rx = (int)q - ry * SRW;
        return gsm0 + 4u * (uint32_t)((sy0 + ry - gby0) * GBW + (sx0 + rx - gbx0));
    };
    float pw[PPT][K];
    uint32_t pa[PPT][K / 2];   // two 16-bit shared-window addresses per register (stage 0)
    uint32_t ov[PPT];          // later pairs: first << 16 | count
    // an unused slot reads its stage's zero word with weight 0: no predicates in the channel loop
    const uint32_t zaddr = gsm0 + (uint32_t)ZERO_OFF;
#pragma unroll
    for (int d = 0; d < PPT; ++d) {
        const int e = (warp * 2 + (d >> 1)) * TW + (d & 1) * 32 + lane;
        const int n = cnt[e];
        ov[d] = n > K ? ((uint32_t)obase[e] << 16) | (uint32_t)(n - K) : 0u;
#pragma unroll
        for (int j = 0; j < K; ++j) {
            float w = 0.0f;
            uint32_t a = zaddr;
            if (j < n) {
                w = ellw[j * NE + e];
                a = box_addr(ellq[j * NE + e]);
            }
            pw[d][j] = w;
            if (j & 1) pa[d][j >> 1] |= a << 16;
            else pa[d][j >> 1] = a;
        }
    }
    // later pairs sorted by element (the order inside an element is the order the atomics are served)
    for (int i = tid; i < n_ovf; i += THREADS) {
        const uint2 pr = tmp[i];
        const int e = (int)(pr.y >> 13);
        const int r = atomicAdd(&cnt2[e], 1);
        ovf[obase[e] + r] = make_uint2(pr.x, box_addr(pr.y & 8191u));
    }
    __syncthreads();  // the build scratch is dead: the load stages may be overwritten
    if (tid < NS) *reinterpret_cast<float*>(smem_raw + tid * G_BYTES + ZERO_OFF) = 0.0f;
    __syncthreads();

    // ---- channel loop
    const int gchunks = (gbh + ROWCHUNK - 1) / ROWCHUNK;
    const int ichunks = NEED_GFLOW ? (bh + ROWCHUNK - 1) / ROWCHUNK : 0;
    const uint32_t full0 = tma::smem_u32(full_bar), ovf0 = gsm0 + (uint32_t)OVF_OFF;
    const uint32_t tx_bytes = (uint32_t)(gchunks * ROWCHUNK * GBW + ichunks * ROWCHUNK * IBW) * 4u;
    auto issue_loads = [&](int i) {  // channel i of the range into stage i % NS (one thread)
        const uint32_t s = (uint32_t)(i % NS);
        const uint32_t bar = full0 + 8u * s;
        if (tx_bytes == 0) { tma::mbar_arrive(bar); return; }
        tma::mbar_arrive_expect_tx(bar, tx_bytes);
        for (int k = 0; k < gchunks; ++k)
            tma::load_3d(gsm0 + s * (uint32_t)G_BYTES + (uint32_t)(k * ROWCHUNK * GBW * 4), &tm_gout, gbx0,
                         gby0 + k * ROWCHUNK, plane0 + i, bar);
        for (int k = 0; k < ichunks; ++k)
            tma::load_3d(gsm0 + (uint32_t)IN_OFF + s * (uint32_t)IN_BYTES + (uint32_t)(k * ROWCHUNK * IBW * 4), &tm_in,
                         bx0, by0 + k * ROWCHUNK, plane0 + i, bar);
    };
    if (tid == 0) {
        fence_async_smem();  // generic-proxy writes to the stages (build scratch) before the async-proxy loads
        for (int i = 0; i < NS && i < nch; ++i) issue_loads(i);
    }

    // grad_flow state: north-west tap address in the input box of stage 0, fractional 1-D weights
    // (wx0 = 1 - wx1 is the same fp32 number as ATen's (x0 + 1) - ix: both differences are exact for
    // x0 >= 1 and the same expression for x0 = 0), east / south steps as bits, running sums
    uint32_t a_n[PPT], tapbits = 0;  // bit k: east tap inside, bit 4 + k: south tap inside, bit 8 + k: pixel inside
    float wx1[PPT], wy1[PPT], gix[PPT], giy[PPT], gcur[PPT];
    const size_t pix0 = (size_t)(ty0 + warp * 2) * p.W + tx0 + lane;  // own pixel k: + (k >> 1) * W + (k & 1) * 32
#pragma unroll
    for (int k = 0; k < PPT; ++k) {
        tapbits |= valid[k] ? 1u << (8 + k) : 0u;
        if (NEED_GFLOW) {
            const int e_nw = (bc[k].t.y0 - by0) * IBW + (bc[k].t.x0 - bx0);
            a_n[k] = gsm0 + (uint32_t)IN_OFF + 4u * (uint32_t)(valid[k] ? e_nw : 0);
            tapbits |= (valid[k] && bc[k].t.x1ok ? 1u << k : 0u) | (valid[k] && bc[k].t.y1ok ? 1u << (4 + k) : 0u);
            wx1[k] = bc[k].wx1;
            wy1[k] = bc[k].wy1;
            gix[k] = giy[k] = 0.0f;
            gcur[k] = valid[k] ? ldg_stream(gout + (size_t)plane0 * plane + pix0 + (size_t)(k >> 1) * p.W + (k & 1) * 32) : 0.0f;
        }
    }
    float* gi_ptr = gin + (size_t)plane0 * plane + pix0;
    const float* go_ptr = gout + (size_t)(plane0 + 1) * plane + pix0;  // own pixels, next channel

    // one loop body for all stages (stage offsets are run-time values: the unrolled-by-stage body
    // was 3 x ~450 instructions and 25 % of the stall samples were instruction fetches)
#pragma unroll 1
    for (int i = 0; i < nch; ++i) {
        const uint32_t st = (uint32_t)(i % NS);
        const uint32_t goff = st * (uint32_t)G_BYTES, ioff = st * (uint32_t)IN_BYTES;
        // (opaque to the optimiser: unpacked addresses / steps are re-derived per channel instead
        // of being hoisted into ~50 more registers)
        asm volatile("" : "+r"(tapbits));
        tma::mbar_wait(full0 + 8u * st, (uint32_t)(i / NS) & 1u);
        if (NEED_GFLOW) {
#pragma unroll
            for (int k = 0; k < PPT; ++k) {
                const uint32_t dx = (tapbits >> k) & 1u ? 4u : 0u, dy = (tapbits >> (4 + k)) & 1u ? 4u * IBW : 0u;
                const uint32_t an = a_n[k] + ioff;
                const float v_nw = lds(an), v_ne = lds(an + dx);
                const float v_sw = lds(an + dy), v_se = lds(an + dy + dx);
                // a tap outside the image re-reads its in-image neighbour with a zero 1-D weight
                const float wx0 = __fsub_rn(1.0f, wx1[k]), wy0 = __fsub_rn(1.0f, wy1[k]);
                const float tx = fmaf(wy1[k], v_se - v_sw, wy0 * (v_ne - v_nw));
                const float ty = fmaf(wx1[k], v_se - v_ne, wx0 * (v_sw - v_nw));
                gix[k] = fmaf(tx, gcur[k], gix[k]);
                giy[k] = fmaf(ty, gcur[k], giy[k]);
                // grad_out of the own pixel for the next channel: in flight during the gathers below
                gcur[k] = ((tapbits >> (8 + k)) & 1u) && i + 1 < nch
                              ? ldg_stream(go_ptr + (size_t)(k >> 1) * p.W + (k & 1) * 32) : 0.0f;
            }
            go_ptr += plane;
        }
        // grad_input: every element's list, summed in list order, stored once.  The stage offset is
        // added to both 16-bit halves of a packed address pair at once (no carry: < 64 KB).
        const uint32_t goff2 = goff * 0x10001u;
        float acc[PPT];
#pragma unroll
        for (int d = 0; d < PPT; ++d) {
            float v[K];
#pragma unroll
            for (int h = 0; h < K / 2; ++h) {
                uint32_t pk = pa[d][h];
                asm volatile("" : "+r"(pk));
                pk += goff2;
                v[2 * h] = lds(pk & 0xffffu);
                v[2 * h + 1] = lds(pk >> 16);
            }
            float a = 0.0f;
#pragma unroll
            for (int j = 0; j < K; ++j) a = fmaf(pw[d][j], v[j], a);
            acc[d] = a;
        }
        // compressive flows / border pile-ups (~3 % of the pairs): the owner walks its elements'
        // later pairs.  (Measured alternatives at 1080p: a warp-cooperative pass with
        // red.shared.add.f32 -- a CAS loop on sm_100a that spun ~5 x per call: 1 795 us; the same with
        // a segmented warp scan and plain read-modify-writes: 1 018 us; this loop not unrolled:
        // 1 064 us; unrolled by the compiler: 921 us.  An L2 prefetch of the boxes 2..8 channels
        // ahead (cp.async.bulk.prefetch.tensor) changed nothing: the waits are not DRAM latency.)
#pragma unroll
        for (int d = 0; d < PPT; ++d) {
            if (ov[d]) {
                uint32_t q = ovf0 + 8u * (ov[d] >> 16);
                for (uint32_t n = ov[d] & 0xffffu; n > 0; --n, q += 8u) {
                    float w;
                    uint32_t a;
                    asm volatile("ld.shared.v2.b32 {%0, %1}, [%2];" : "=f"(w), "=r"(a) : "r"(q));
                    acc[d] = fmaf(w, lds(a + goff), acc[d]);
                }
            }
        }
#pragma unroll
        for (int d = 0; d < PPT; ++d)
            if ((tapbits >> (8 + d)) & 1u) st_stream1(gi_ptr + (size_t)(d >> 1) * p.W + (d & 1) * 32, acc[d]);
        gi_ptr += plane;
        __syncthreads();  // stage consumed by every thread
        if (tid == 0 && i + NS < nch) issue_loads(i + NS);
    }
    if (NEED_GFLOW) {
#pragma unroll
        for (int k = 0; k < PPT; ++k) {
            if (!((tapbits >> (8 + k)) & 1u)) continue;
            const int x = tx0 + (k & 1) * 32 + lane, y = ty0 + warp * 2 + (k >> 1);
            const size_t pix = (size_t)y * p.W + x;
            // (recomputed: fewer live registers in the loop)
            const BwdCoord c2 = bwd_coord(__ldg(lin_x + x), __ldg(lin_y + y), __ldg(fl + pix), __ldg(fl + plane + pix), p);
            store_gflow(gflow, p, b, pix, c2, gix[k], giy[k], acc_gflow);
        }
    }
}

// Fix-up: one thread per output pixel.  A tap is missing from the main launch's result iff its
// destination tile was flagged (boxes / lists did not fit) or the pixel lies outside that tile's
// search region; exactly those taps are added with global float reductions.
__global__ void __launch_bounds__(256)
warp_bwd_fixup_kernel(const float* __restrict__ gout, const float* __restrict__ flow, float* __restrict__ gin,
                      const float* __restrict__ lin_x, const float* __restrict__ lin_y, WarpParams p,
                      int tiles_x, int tiles_y, const unsigned char* __restrict__ flags) {
    using namespace bg;
    const int x = blockIdx.x * 32 + threadIdx.x, y = blockIdx.y * 8 + threadIdx.y, b = blockIdx.z;
    if (x >= p.W || y >= p.H) return;
    const size_t plane = (size_t)p.H * p.W;
    const size_t pix = (size_t)y * p.W + x;
    const float* fl = flow + (size_t)b * 2 * plane;
    const BwdCoord bc = bwd_coord(__ldg(lin_x + x), __ldg(lin_y + y), __ldg(fl + pix), __ldg(fl + plane + pix), p);
    const Taps& t = bc.t;
    auto missing = [&](int X, int Y) -> bool {
        const int txi = X / TW, tyi = Y / TH;
        if (flags[(b * tiles_y + tyi) * tiles_x + txi]) return true;
        int sx0, sy0;
        search_origin(fl, plane, p, txi * TW, tyi * TH, sx0, sy0);
        return x < sx0 || x >= sx0 + SRW || y < sy0 || y >= sy0 + SRH;
    };
    const bool m_nw = missing(t.x0, t.y0);
    const bool m_ne = t.x1ok && missing(t.x0 + 1, t.y0);
    const bool m_sw = t.y1ok && missing(t.x0, t.y0 + 1);
    const bool m_se = t.x1ok && t.y1ok && missing(t.x0 + 1, t.y0 + 1);
    if (!(m_nw || m_ne || m_sw || m_se)) return;
    const float* gp = gout + (size_t)b * p.C * plane + pix;
    float* gi = gin + (size_t)b * p.C * plane + (size_t)t.y0 * p.W + t.x0;
#pragma unroll 4
    for (int c = 0; c < p.C; ++c) {
        const float g = __ldg(gp);
        if (m_nw) atomicAdd(gi, __fmul_rn(t.nw, g));
        if (m_ne) atomicAdd(gi + 1, __fmul_rn(t.ne, g));
        if (m_sw) atomicAdd(gi + p.W, __fmul_rn(t.sw, g));
        if (m_se) atomicAdd(gi + p.W + 1, __fmul_rn(t.se, g));
        gp += plane;
        gi += plane;
    }
}

}  // namespace dsvc

using namespace dsvc;

static bool encode_xy_plane_box(CUtensorMap* tm, const float* base, const WarpParams& p, int box_w, int box_h) {
    auto encode = tensor_map_encoder();
    if (!encode) return false;
    const cuuint64_t gdim[3] = {(cuuint64_t)p.W, (cuuint64_t)p.H, (cuuint64_t)p.B * p.C};
    const cuuint64_t gstride[2] = {(cuuint64_t)p.W * 4, (cuuint64_t)p.H * p.W * 4};
    const cuuint32_t box[3] = {(cuuint32_t)box_w, (cuuint32_t)box_h, 1};
    const cuuint32_t estr[3] = {1, 1, 1};
    return encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, const_cast<float*>(base), gdim, gstride, box, estr,
                  CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

size_t dsvc_warp_bwd_gather_workspace(int B, int H, int W) {
    const long long tiles = (long long)((W + bg::TW - 1) / bg::TW) * ((H + bg::TH - 1) / bg::TH) * B;
    return (size_t)((tiles + 255) / 256 * 256);
}

// returns -1 when the shape is not eligible (the caller uses another kernel).  grad_input needs
// no initialisation (every element is written); grad_flow may be uninitialised.  `workspace`:
// dsvc_warp_bwd_gather_workspace() bytes (one flag per tile, written before it is read).
int dsvc_warp_bwd_gather_launch(const float* gout, const float* input, const float* flow, float* gin,
                                float* gflow, const float* lin_x, const float* lin_y, const WarpParams& p,
                                bool force, void* workspace, size_t workspace_bytes, cudaStream_t st) {
    if (!gin) return -1;  // flow-only gradients: the other kernels
    if (p.W % 4 != 0 || !aligned16(input) || !aligned16(gout)) return -1;
    if (!force && (p.C < 8 || p.W < 64 || p.H < 16)) return -1;
    if ((long long)p.B * p.C > (1ll << 30)) return -1;
    if (!workspace || workspace_bytes < dsvc_warp_bwd_gather_workspace(p.B, p.H, p.W)) return -1;
    CUtensorMap tm_in, tm_gout;
    if (!encode_xy_plane_box(&tm_in, input, p, bg::IBW, bg::ROWCHUNK)) return -1;
    if (!encode_xy_plane_box(&tm_gout, gout, p, bg::GBW, bg::ROWCHUNK)) return -1;
    static unsigned long long attr_set = 0;
    static int sms_of[64];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return (int)cudaErrorInvalidDevice;
    if (!((attr_set >> dev) & 1ull)) {
        cudaError_t e = cudaFuncSetAttribute(warp_bwd_gather_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bg::SMEM_BYTES);
        if (e == cudaSuccess)
            e = cudaFuncSetAttribute(warp_bwd_gather_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)bg::SMEM_BYTES);
        if (e != cudaSuccess) return (int)e;
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = DSVC_NUM_SMS;
        sms_of[dev] = n;
        attr_set |= 1ull << dev;
    }
    const int tiles_x = (p.W + bg::TW - 1) / bg::TW, tiles_y = (p.H + bg::TH - 1) / bg::TH;
    const long long ntiles = (long long)tiles_x * tiles_y * p.B;
    if (ntiles > (1ll << 24)) return -1;
    // fewer tiles than resident CTAs: cut the channels in ranges (the per-tile list build is
    // repeated per range)
    const int slots = 2 * sms_of[dev];
    int csplit = 1;
    while (ntiles * csplit < (long long)slots && p.C / (csplit * 2) >= 16) csplit *= 2;
    const int cper = (p.C + csplit - 1) / csplit;
    if (gflow && csplit > 1) {
        const cudaError_t e = cudaMemsetAsync(gflow, 0, (size_t)p.B * 2 * p.H * p.W * sizeof(float), st);
        if (e != cudaSuccess) return (int)e;
    }
    unsigned char* flags = static_cast<unsigned char*>(workspace);
    const unsigned grid = (unsigned)(ntiles * csplit);
    if (gflow)
        warp_bwd_gather_kernel<true><<<grid, bg::THREADS, bg::SMEM_BYTES, st>>>(
            tm_in, tm_gout, gout, input, flow, gin, gflow, lin_x, lin_y, p, tiles_x, tiles_y, csplit, cper, flags);
    else
        warp_bwd_gather_kernel<false><<<grid, bg::THREADS, bg::SMEM_BYTES, st>>>(
            tm_in, tm_gout, gout, input, flow, gin, gflow, lin_x, lin_y, p, tiles_x, tiles_y, csplit, cper, flags);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) return (int)e;
    dim3 block(32, 8), fgrid((p.W + 31) / 32, (p.H + 7) / 8, p.B);
    warp_bwd_fixup_kernel<<<fgrid, block, 0, st>>>(gout, flow, gin, lin_x, lin_y, p, tiles_x, tiles_y, flags);
    return (int)cudaGetLastError();
}
