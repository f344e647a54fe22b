// tcgen05 + TMA implicit-GEMM convolution, stride 1, with the activation HALO TILE staged once per
// 64-channel chunk and re-used by every filter tap.
//
//   D[pixel][co] = sum_{chunk, tap, c} A[pixel + tap][chunk*64 + c] * W[tap][co][chunk*64 + c]
//
// The per-tap kernel in tc_conv.cu re-reads the 128-pixel activation box from L2 for every tap (9x for a
// 3x3, 49x for a 7x7 filter) and is L2->smem-fill bound (ncu: tensor pipe 64 % of active cycles, see
// profiles/).  Here a CTA owns an 8 (x) by 16 (y) pixel tile and loads the (8+kw-1) x (16+kh-1) halo
// tile of one 64-channel chunk with ONE TMA box (hi and lo planes); the UMMA A descriptor of tap (ky,kx)
// simply starts (ky*pitch + kx) rows further into that tile with SBO = pitch*128 B, because an 8-pixel row
// of the output tile is 8 consecutive 128-byte smem rows (the SWIZZLE_128B XOR acts on absolute smem
// address bits, so row-granular start offsets read back exactly what TMA wrote).  Weights stream
// through their own ring of (tap, chunk) stages (BN rows x 64 channels, hi + lo, SWIZZLE_128B — a first
// version with 32-channel SWIZZLE_64B stages was correct but slower: 64-byte rows halve the useful
// bytes per shared-memory wavefront of the B operand).  fp32 fidelity: 3 bf16 MMAs per product
// (A_lo*W_hi + A_hi*W_lo + A_hi*W_hi), fp32 accumulation in TMEM.
//
// Warp roles (128 threads): warp 0 lane 0 = weight TMA producer, warp 2 lane 0 = activation TMA
// producer, warp 1 = TMEM allocator + (lane 0) MMA issuer, all four warps = epilogue.
// Channel counts: ci any multiple of 8 (TMA zero-fills the box beyond the tensor, the K loop only issues
// the 16-channel steps that exist), co any value with N tile in {16, 64, 128, 256} (rows beyond co are
// zero-filled by TMA and never stored).
#include <cstdlib>
#include "tc_common.cuh"

namespace skit {
namespace tc {

// One launch may cover several rectangular REGIONS of the output, each with its own sub-filter (a rectangle of taps of the
// packed filter), operand origin and halo-box shape.  A plain convolution is one region.  The stride-1 input gradient — a full
// correlation of the zero-haloed output gradient — is an interior region with all taps plus four one-pixel border strips whose
// windows only reach the data through one filter row / column: the strips cost a third of a tile and fill the SMs the interior
// leaves idle, instead of pushing a 130x130 map to 153 tiles = two waves on 148 SMs.
struct HaloRegion {
    int first;                       // first CTA (blockIdx.x) of this region
    int kh, kw;                      // sub-filter taps
    int tap_base, tap_sy, tap_sx;    // tap index in the packed filter = tap_base + ky*tap_sy + kx*tap_sx
    int org_y, org_x;                // operand coordinate read by output (0,0) of the region through sub-filter tap (0,0)
    int ho, wo, tiles_x;             // region size in output pixels, tiles per row
    int ooy, oox;                    // where the region's output (0,0) lands in the output map (before osy/osx scaling)
    int amap;                        // which activation tensor-map pair (box shape) it uses
    int pitch, a_rows;               // halo tile width in pixels = 8 + kw - 1; rows*cols of the halo tile
};
constexpr int MAX_REGIONS = 5;
constexpr int MAX_AMAPS = 3;

struct AMaps {
    CUtensorMap hi[MAX_AMAPS];
    CUtensorMap lo[MAX_AMAPS];
};

struct TcHaloP {
    int nreg;
    HaloRegion reg[MAX_REGIONS];
    int kc;            // 64-channel chunks (ceil(ci/64))
    int kk_last;       // 16-channel K steps in the last chunk (1..4)
    int co;
    int OH, OW, osy, osx, poy, pox;   // output map, placement stride and phase offset (stride-2 dgrad writes interleaved quarters)
    const float* bias;
    float* y;
    double* stats;
    int stats_per_n;
    int a_plane;       // bytes reserved per A plane (largest region's a_rows*128 rounded up to 1024)
    int nw;            // weight stages
    int na;            // activation stages (2, or 1 when a large halo tile must leave room for wide weight stages)
    int terms;         // 3: A_lo*W_hi + A_hi*W_lo + A_hi*W_hi (fp32-parity forward);  2: (A_hi + A_lo)*W_hi only — the W_lo plane is
                       // neither loaded nor multiplied (input-gradient launches, skit_set_backward_terms)
    long long* dbg;    // optional per-CTA clock64 stamps [cta][8] (skit_debug_set_buffer), NULL in production
    int dbg_mode;      // limiter experiments (SKIT_DBG_MODE, tools/bench_limiter.py; results are wrong): 1 = no MMAs issued (load pipeline
                       // alone), 2 = no filter TMA (MMA pipeline alone, stale filter), 3 = no activation TMA
};

constexpr int NA_MAX = 2;  // activation stages

template <int BN>
__global__ void __launch_bounds__(128, 1)
conv_tc_halo_kernel(const __grid_constant__ AMaps tmA,
                    const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo, const TcHaloP p) {
    pdl_trigger();       // PDL: the next kernel in the stream may be scheduled once every CTA of this grid has started
    constexpr int W_PLANE = BN * 128;       // BN rows x 64 channels x 2 B
    // a two-term launch stages the hi plane only: half-size stages, twice as many of them — the bytes in flight per SM (stages x
    // stage size against ~1.7 us of L2 latency under load), not the MMA rate, bound the two-term kernel with two 64 KB stages
    const int W_STAGE = p.terms == 2 ? W_PLANE : 2 * W_PLANE;
    constexpr int TCOLS = BN < 32 ? 32 : BN;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
    const int a_stage = 2 * p.a_plane;
    const int NA = p.na;
    const uint32_t w0 = smem0 + NA * a_stage;
    const uint32_t bar0 = w0 + p.nw * W_STAGE;
    auto a_full = [&](int s) { return bar0 + 8u * s; };
    auto a_empty = [&](int s) { return bar0 + 8u * (NA + s); };
    auto w_full = [&](int s) { return bar0 + 8u * (2 * NA + s); };
    auto w_empty = [&](int s) { return bar0 + 8u * (2 * NA + p.nw + s); };
    const uint32_t tmem_full_bar = bar0 + 8u * (2 * NA + 2 * p.nw);
    const uint32_t tmem_slot = tmem_full_bar + 8u;
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem0));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.z;
    const int n0 = blockIdx.y * BN;
    int ri = 0;
    for (int r = 1; r < p.nreg; r++)
        if ((int)blockIdx.x >= p.reg[r].first) ri = r;
    // the region's fields go to registers once: a reference into the parameter array with a runtime index would be re-read
    // through indexed constant loads (or a local copy) inside the epilogue's store loop
    HaloRegion R;
    {
        const int* src = reinterpret_cast<const int*>(&p.reg[0]);
        int* dst = reinterpret_cast<int*>(&R);
#pragma unroll
        for (int i = 0; i < (int)(sizeof(HaloRegion) / sizeof(int)); i++) {
            int v = src[i];
#pragma unroll
            for (int r = 1; r < MAX_REGIONS; r++)
                if (r == ri) v = src[r * (int)(sizeof(HaloRegion) / sizeof(int)) + i];
            dst[i] = v;
        }
    }
    const int lt = blockIdx.x - R.first;
    const int tile_y = lt / R.tiles_x, tile_x = lt - tile_y * R.tiles_x;
    const int y0 = tile_y * 16, x0 = tile_x * 8;
    const CUtensorMap* tmA_hi = &tmA.hi[R.amap];
    const CUtensorMap* tmA_lo = &tmA.lo[R.amap];
    long long* dbg = p.dbg ? p.dbg + 8ll * (blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z)) : nullptr;
    if (dbg && threadIdx.x == 0) dbg[0] = clock64();

    if (threadIdx.x == 0) {
        tma_prefetch_desc(tmA_hi); tma_prefetch_desc(tmA_lo);
        tma_prefetch_desc(&tmW_hi); tma_prefetch_desc(&tmW_lo);
        for (int s = 0; s < NA; s++) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
        for (int s = 0; s < p.nw; s++) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
        mbar_init(tmem_full_bar, 1);
        mbar_fence_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc<TCOLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();          // PDL: barrier init / TMEM allocation / descriptor prefetch above overlap the predecessor's tail
    const uint32_t tmem_base = *tmem_slot_gen;
    if (dbg && threadIdx.x == 0) dbg[1] = clock64();

    if (warp == 2 && lane == 0) {
        // ---------------- activation producer: one halo box (hi + lo) per 64-channel chunk
        const uint32_t bytes = 2u * (uint32_t)R.a_rows * 128u;
        for (int c = 0; c < p.kc; c++) {
            const int s = c % NA, ph = (c / NA) & 1;
            mbar_wait(a_empty(s), ph ^ 1);
            if (p.dbg_mode == 3) { mbar_arrive(a_full(s)); continue; }
            mbar_expect_tx(a_full(s), bytes);
            const uint32_t sa = smem0 + s * a_stage;
            tma_load_4d(sa, tmA_hi, a_full(s), c * 64, R.org_x + x0, R.org_y + y0, n);
            tma_load_4d(sa + p.a_plane, tmA_lo, a_full(s), c * 64, R.org_x + x0, R.org_y + y0, n);
        }
    } else if (warp == 0 && lane == 0) {
        // ---------------- weight producer: (chunk, tap) stages
        int it = 0;
        for (int c = 0; c < p.kc; c++)
            for (int ky = 0; ky < R.kh; ky++)
                for (int kx = 0; kx < R.kw; kx++, it++) {
                    const int s = it % p.nw, ph = (it / p.nw) & 1;
                    const int tap = R.tap_base + ky * R.tap_sy + kx * R.tap_sx;
                    mbar_wait(w_empty(s), ph ^ 1);
                    if (p.dbg_mode == 2) { mbar_arrive(w_full(s)); continue; }
                    mbar_expect_tx(w_full(s), p.terms == 2 ? W_PLANE : W_STAGE);
                    const uint32_t sw = w0 + s * W_STAGE;
                    tma_load_3d(sw, &tmW_hi, w_full(s), c * 64, n0, tap);
                    if (p.terms != 2) tma_load_3d(sw + W_PLANE, &tmW_lo, w_full(s), c * 64, n0, tap);
                }
    } else if (warp == 1) {
        // ---------------- MMA issuer.  The WHOLE warp walks the loop with warp-uniform control flow and one elected lane issues:
        // under a divergent `lane == 0` branch the compiler wraps every tcgen05.mma in an ELECT / BRA.U.ANY loop with per-use
        // register -> uniform-register moves, and the issue rate — not the tensor pipe — set the pace of the main loop
        // (tools/bench_limiter.py: removing the filter TMA changed nothing, adding branches to the issue path cost 10 %).
        // Descriptors: constant high part + (address >> 4); a K step of 16 bf16 advances the start address by 32 B = 2 units.
        constexpr uint32_t idesc = make_idesc_bf16(BN, 0, 0);
        const uint32_t sbo_a = (uint32_t)R.pitch * 128u;
        const uint64_t dA = make_desc(0, 16, sbo_a), dW = make_desc(0, 16, 1024);
        const bool leader = elect_one();
        const int dmode = p.dbg_mode;
        int it = 0;
        uint32_t acc = 0;
        for (int c = 0; c < p.kc; c++) {
            const int s = c % NA, ph = (c / NA) & 1;
            const int kkc = (c == p.kc - 1) ? p.kk_last : 4;
            mbar_wait(a_full(s), ph);
            tc_fence_after();
            if (dbg && c == 0 && leader) dbg[2] = clock64();
            const uint32_t sa = smem0 + s * a_stage;
            for (int ky = 0; ky < R.kh; ky++)
                for (int kx = 0; kx < R.kw; kx++, it++) {
                    const uint32_t arow = sa + (uint32_t)(ky * R.pitch + kx) * 128u;
                    const int ws = it % p.nw, wph = (it / p.nw) & 1;
                    mbar_wait(w_full(ws), wph);
                    tc_fence_after();
                    const uint32_t sw = w0 + ws * W_STAGE;
                    if (leader) {
                        const uint64_t a_hi = dA + (arow >> 4), a_lo = dA + ((arow + (uint32_t)p.a_plane) >> 4);
                        const uint64_t w_hi = dW + (sw >> 4), w_lo = dW + ((sw + W_PLANE) >> 4);
                        if (dmode == 1) {              // limiter experiment: no MMAs beyond the first
                            if (it == 0) mma_bf16(tmem_base, a_hi, w_hi, idesc, 0u);
                        } else if (p.terms == 2) {
#pragma unroll 4
                            for (int kk = 0; kk < kkc; kk++) {
                                mma_bf16(tmem_base, a_lo + 2 * kk, w_hi + 2 * kk, idesc, acc);
                                acc = 1u;
                                mma_bf16(tmem_base, a_hi + 2 * kk, w_hi + 2 * kk, idesc, 1u);
                            }
                        } else {
#pragma unroll 4
                            for (int kk = 0; kk < kkc; kk++) {
                                mma_bf16(tmem_base, a_lo + 2 * kk, w_hi + 2 * kk, idesc, acc);
                                acc = 1u;
                                mma_bf16(tmem_base, a_hi + 2 * kk, w_lo + 2 * kk, idesc, 1u);
                                mma_bf16(tmem_base, a_hi + 2 * kk, w_hi + 2 * kk, idesc, 1u);
                            }
                        }
                        mma_commit(w_empty(ws));
                    }
                    __syncwarp();
                }
            if (leader) mma_commit(a_empty(s));
            __syncwarp();
        }
        if (leader) {
            mma_commit(tmem_full_bar);
            if (dbg) dbg[3] = clock64();
        }
    }
    __syncwarp();

    // ---------------- epilogue (all 4 warps): TMEM lane r = ty*8 + tx
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    __syncwarp();
    if (dbg && threadIdx.x == 0) dbg[4] = clock64();
    // The stage memory is free now: per-warp staging tile [32 rows][36 floats] (16-byte aligned rows whose stride
    // keeps both the float4 row writes and the float4 row reads bank-conflict free), then the statistics scratch.
    constexpr int STG = 36;
    float* stg = reinterpret_cast<float*>(smem_gen) + warp * (32 * STG);
    float* red = reinterpret_cast<float*>(smem_gen) + 4 * 32 * STG;  // [4 warps][TCOLS][2]
    {
        const int r = warp * 32 + lane;
        const int ty = r >> 3, tx = r & 7;
        const int oy = y0 + ty, ox = x0 + tx;
        const int py = (oy + R.ooy) * p.osy + p.poy, px = (ox + R.oox) * p.osx + p.pox;
        const bool valid = oy < R.ho && ox < R.wo && py < p.OH && px < p.OW;
        const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
        float* yrow = p.y + (((long long)n * p.OH + py) * p.OW + px) * p.co + n0;
        // the 8 output rows this lane stores (4 rows per pass, 8 lanes per 128-byte row): pointers and validity once per tile
        float* rowp[8];
        uint32_t rowok = 0;
#pragma unroll
        for (int it = 0; it < 8; it++) {
            const int rr = it * 4 + (lane >> 3);
            const int g = warp * 32 + rr;
            const int gy = (y0 + (g >> 3) + R.ooy) * p.osy + p.poy, gx = (x0 + (g & 7) + R.oox) * p.osx + p.pox;
            rowp[it] = p.y + (((long long)n * p.OH + gy) * p.OW + gx) * p.co + n0 + (lane & 7) * 4;
            rowok |= ((vmask >> rr) & 1u) << it;
        }
#pragma unroll 1
        for (int c = 0; c < TCOLS; c += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
            if (BN >= 32 && n0 + c + 32 <= p.co) {
                if (p.bias) {
#pragma unroll
                    for (int j = 0; j < 32; j++) v[j] += __ldg(p.bias + n0 + c + j);
                }
                // registers (one pixel row per lane) -> smem tile -> coalesced 128-byte rows in global memory
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(stg + lane * STG + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                __syncwarp();
                float4 t4[8];
#pragma unroll
                for (int it = 0; it < 8; it++)      // 8 independent 16-byte shared loads first ...
                    t4[it] = *reinterpret_cast<const float4*>(stg + (it * 4 + (lane >> 3)) * STG + (lane & 7) * 4);
#pragma unroll
                for (int it = 0; it < 8; it++)      // ... then the 8 coalesced row stores
                    if ((rowok >> it) & 1u) *reinterpret_cast<float4*>(rowp[it] + c) = t4[it];
                if (p.stats) {   // lane j owns column c + j: sum it down the 32 staged rows
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int rr = 0; rr < 32; rr++) {
                        const float x = ((vmask >> rr) & 1u) ? stg[rr * STG + lane] : 0.f;
                        s1 += x; s2 = fmaf(x, x, s2);
                    }
                    red[(warp * TCOLS + c + lane) * 2 + 0] = s1;
                    red[(warp * TCOLS + c + lane) * 2 + 1] = s2;
                }
                __syncwarp();
            } else {  // ragged channel tail (co not a multiple of 32, e.g. the 5-channel generator head)
#pragma unroll
                for (int j = 0; j < 32; j++) {
                    const bool cv = n0 + c + j < p.co;
                    v[j] = cv ? v[j] + (p.bias ? __ldg(p.bias + n0 + c + j) : 0.f) : 0.f;
                    if (valid && cv) yrow[c + j] = v[j];
                }
                if (p.stats) {
                    float sq[32];
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        v[j] = valid ? v[j] : 0.f;
                        sq[j] = v[j] * v[j];
                    }
                    const float s1 = col_reduce32(v, lane);
                    const float s2 = col_reduce32(sq, lane);
                    red[(warp * TCOLS + c + lane) * 2 + 0] = s1;
                    red[(warp * TCOLS + c + lane) * 2 + 1] = s2;
                }
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (dbg && threadIdx.x == 0) dbg[5] = clock64();
    if (p.stats) {
        for (int col = threadIdx.x; col < TCOLS; col += 128) {
            if (n0 + col >= p.co) continue;
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int w = 0; w < 4; w++) { s1 += red[(w * TCOLS + col) * 2]; s2 += red[(w * TCOLS + col) * 2 + 1]; }
            double* dst = p.stats + ((long long)(p.stats_per_n ? n : 0) * p.co + n0 + col) * 2;
            atomicAdd(dst, (double)s1);
            atomicAdd(dst + 1, (double)s2);
        }
    }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TCOLS>(tmem_base);
    }
    if (dbg && threadIdx.x == 0) dbg[6] = clock64();
}

// ---------------------------------------------------------------------------------------------------------------------
// Persistent variant for layers that run several waves with N tiles of 128 or less (64 -> 128 / 128 -> 64 convs at full
// resolution and their input gradients): one CTA per SM walks a strided list of (image, N tile, pixel tile) units.  Two TMEM
// accumulators alternate, so the epilogue of unit i (dedicated warps 4..7: TMEM -> registers -> smem staging -> coalesced
// stores + statistics) overlaps the main loop of unit i+1, and barrier / TMEM / descriptor setup is paid once per SM instead of
// once per tile.  For these layers the K loop is short (9..18 tap steps), so prologue + epilogue were ~40 % of a tile.
// Warp roles (256 threads): 0 = weight TMA, 1 = MMA issuer + TMEM allocator, 2 = activation TMA, 3 idle, 4..7 = epilogue.
struct TcPersistP {
    TcHaloP h;            // single region in h.reg[0]
    int tiles, ntiles, n_img;   // pixel tiles per image, N tiles, images
    int units;            // tiles * ntiles * n_img
};

template <int BN>
__global__ void __launch_bounds__(256, 1)
conv_tc_halo_persist_kernel(const __grid_constant__ AMaps tmA, const __grid_constant__ CUtensorMap tmW_hi,
                            const __grid_constant__ CUtensorMap tmW_lo, const TcPersistP pp) {
    pdl_trigger();       // PDL: the next kernel in the stream may be scheduled once every CTA of this grid has started
    constexpr int W_PLANE = BN * 128;
    constexpr int TCOLS = BN < 32 ? 32 : BN;
    constexpr int STG = 36;
    constexpr int EPI_BYTES = 4 * 32 * STG * 4 + 4 * TCOLS * 2 * 4;   // staging tiles + statistics scratch of the 4 epilogue warps
    const TcHaloP& p = pp.h;
    const int W_STAGE = p.terms == 2 ? W_PLANE : 2 * W_PLANE;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
    const int a_stage = 2 * p.a_plane;
    const int NA = p.na;
    const uint32_t w0 = smem0 + NA * a_stage;
    const uint32_t epi0 = w0 + p.nw * W_STAGE;
    const uint32_t bar0 = epi0 + EPI_BYTES;
    auto a_full = [&](int s) { return bar0 + 8u * s; };
    auto a_empty = [&](int s) { return bar0 + 8u * (NA + s); };
    auto w_full = [&](int s) { return bar0 + 8u * (2 * NA + s); };
    auto w_empty = [&](int s) { return bar0 + 8u * (2 * NA + p.nw + s); };
    auto acc_full = [&](int b) { return bar0 + 8u * (2 * NA + 2 * p.nw + b); };
    auto acc_empty = [&](int b) { return bar0 + 8u * (2 * NA + 2 * p.nw + 2 + b); };
    const uint32_t tmem_slot = bar0 + 8u * (2 * NA + 2 * p.nw + 4);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem0));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const HaloRegion R = p.reg[0];
    const CUtensorMap* tmA_hi = &tmA.hi[0];
    const CUtensorMap* tmA_lo = &tmA.lo[0];

    if (threadIdx.x == 0) {
        tma_prefetch_desc(tmA_hi); tma_prefetch_desc(tmA_lo);
        tma_prefetch_desc(&tmW_hi); tma_prefetch_desc(&tmW_lo);
        for (int s = 0; s < NA; s++) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
        for (int s = 0; s < p.nw; s++) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
        for (int b = 0; b < 2; b++) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 4); }
        mbar_fence_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc<2 * TCOLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();          // PDL: barrier init / TMEM allocation / descriptor prefetch above overlap the predecessor's tail
    const uint32_t tmem_base = *tmem_slot_gen;
    const int tiles_per_n = pp.tiles * pp.ntiles;

    // unit u -> (image, N tile, pixel tile): pixel tiles fastest, so neighbouring CTAs share the weight tile in L2
    auto decode = [&](int u, int& n, int& n0, int& y0, int& x0) {
        n = u / tiles_per_n;
        const int r = u - n * tiles_per_n;
        const int nt = r / pp.tiles, t = r - nt * pp.tiles;
        n0 = nt * BN;
        const int ty = t / R.tiles_x;
        y0 = ty * 16; x0 = (t - ty * R.tiles_x) * 8;
    };

    if (warp == 2 && lane == 0) {
        // ---------------- activation producer
        const uint32_t bytes = 2u * (uint32_t)R.a_rows * 128u;
        int it = 0;
        for (int u = blockIdx.x; u < pp.units; u += gridDim.x) {
            int n, n0, y0, x0;
            decode(u, n, n0, y0, x0);
            for (int c = 0; c < p.kc; c++, it++) {
                const int s = it % NA, ph = (it / NA) & 1;
                mbar_wait(a_empty(s), ph ^ 1);
                mbar_expect_tx(a_full(s), bytes);
                const uint32_t sa = smem0 + s * a_stage;
                tma_load_4d(sa, tmA_hi, a_full(s), c * 64, R.org_x + x0, R.org_y + y0, n);
                tma_load_4d(sa + p.a_plane, tmA_lo, a_full(s), c * 64, R.org_x + x0, R.org_y + y0, n);
            }
        }
    } else if (warp == 0 && lane == 0) {
        // ---------------- weight producer
        int it = 0;
        for (int u = blockIdx.x; u < pp.units; u += gridDim.x) {
            int n, n0, y0, x0;
            decode(u, n, n0, y0, x0);
            for (int c = 0; c < p.kc; c++)
                for (int ky = 0; ky < R.kh; ky++)
                    for (int kx = 0; kx < R.kw; kx++, it++) {
                        const int s = it % p.nw, ph = (it / p.nw) & 1;
                        const int tap = R.tap_base + ky * R.tap_sy + kx * R.tap_sx;
                        mbar_wait(w_empty(s), ph ^ 1);
                        mbar_expect_tx(w_full(s), p.terms == 2 ? W_PLANE : W_STAGE);
                        const uint32_t sw = w0 + s * W_STAGE;
                        tma_load_3d(sw, &tmW_hi, w_full(s), c * 64, n0, tap);
                        if (p.terms != 2) tma_load_3d(sw + W_PLANE, &tmW_lo, w_full(s), c * 64, n0, tap);
                    }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: accumulator (unit index & 1).  Whole warp, warp-uniform control flow, one elected lane issues
        // (see conv_tc_halo_kernel)
        constexpr uint32_t idesc = make_idesc_bf16(BN, 0, 0);
        const uint32_t sbo_a = (uint32_t)R.pitch * 128u;
        const uint64_t dA = make_desc(0, 16, sbo_a), dW = make_desc(0, 16, 1024);
        const bool leader = elect_one();
        int ita = 0, itw = 0, ui = 0;
        for (int u = blockIdx.x; u < pp.units; u += gridDim.x, ui++) {
            const int b = ui & 1;
            mbar_wait(acc_empty(b), ((ui >> 1) & 1) ^ 1);     // the epilogue has drained this accumulator (first two uses pass)
            tc_fence_after();
            const uint32_t dacc = tmem_base + (uint32_t)(b * TCOLS);
            uint32_t acc = 0;
            for (int c = 0; c < p.kc; c++, ita++) {
                const int s = ita % NA, ph = (ita / NA) & 1;
                const int kkc = (c == p.kc - 1) ? p.kk_last : 4;
                mbar_wait(a_full(s), ph);
                tc_fence_after();
                const uint32_t sa = smem0 + s * a_stage;
                for (int ky = 0; ky < R.kh; ky++)
                    for (int kx = 0; kx < R.kw; kx++, itw++) {
                        const uint32_t arow = sa + (uint32_t)(ky * R.pitch + kx) * 128u;
                        const int ws = itw % p.nw, wph = (itw / p.nw) & 1;
                        mbar_wait(w_full(ws), wph);
                        tc_fence_after();
                        const uint32_t sw = w0 + ws * W_STAGE;
                        if (leader) {
                            const uint64_t a_hi = dA + (arow >> 4), a_lo = dA + ((arow + (uint32_t)p.a_plane) >> 4);
                            const uint64_t w_hi = dW + (sw >> 4), w_lo = dW + ((sw + W_PLANE) >> 4);
                            if (p.terms == 2) {
#pragma unroll 4
                                for (int kk = 0; kk < kkc; kk++) {
                                    mma_bf16(dacc, a_lo + 2 * kk, w_hi + 2 * kk, idesc, acc);
                                    acc = 1u;
                                    mma_bf16(dacc, a_hi + 2 * kk, w_hi + 2 * kk, idesc, 1u);
                                }
                            } else {
#pragma unroll 4
                                for (int kk = 0; kk < kkc; kk++) {
                                    mma_bf16(dacc, a_lo + 2 * kk, w_hi + 2 * kk, idesc, acc);
                                    acc = 1u;
                                    mma_bf16(dacc, a_hi + 2 * kk, w_lo + 2 * kk, idesc, 1u);
                                    mma_bf16(dacc, a_hi + 2 * kk, w_hi + 2 * kk, idesc, 1u);
                                }
                            }
                            mma_commit(w_empty(ws));
                        }
                        __syncwarp();
                    }
                if (leader) mma_commit(a_empty(s));
                __syncwarp();
            }
            if (leader) mma_commit(acc_full(b));
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ---------------- epilogue warps: TMEM lanes 32*(warp-4) .. +31, pixel r = ty*8 + tx of the unit's tile
        const int ew = warp - 4;
        float* stg = reinterpret_cast<float*>(smem_gen + (epi0 - smem0)) + ew * (32 * STG);
        float* red = reinterpret_cast<float*>(smem_gen + (epi0 - smem0)) + 4 * 32 * STG;   // [4][TCOLS][2]
        int ui = 0;
        for (int u = blockIdx.x; u < pp.units; u += gridDim.x, ui++) {
            int n, n0, y0, x0;
            decode(u, n, n0, y0, x0);
            const int b = ui & 1;
            mbar_wait(acc_full(b), (ui >> 1) & 1);
            tc_fence_after();
            const int r = ew * 32 + lane;
            const int oy = y0 + (r >> 3), ox = x0 + (r & 7);
            const int py = (oy + R.ooy) * p.osy + p.poy, px = (ox + R.oox) * p.osx + p.pox;
            const bool valid = oy < R.ho && ox < R.wo && py < p.OH && px < p.OW;
            const uint32_t vmask = __ballot_sync(0xffffffffu, valid);
            float* yrow = p.y + (((long long)n * p.OH + py) * p.OW + px) * p.co + n0;
            float* rowp[8];
            uint32_t rowok = 0;
#pragma unroll
            for (int it = 0; it < 8; it++) {
                const int rr = it * 4 + (lane >> 3);
                const int g = ew * 32 + rr;
                const int gy = (y0 + (g >> 3) + R.ooy) * p.osy + p.poy, gx = (x0 + (g & 7) + R.oox) * p.osx + p.pox;
                rowp[it] = p.y + (((long long)n * p.OH + gy) * p.OW + gx) * p.co + n0 + (lane & 7) * 4;
                rowok |= ((vmask >> rr) & 1u) << it;
            }
#pragma unroll 1
            for (int c = 0; c < TCOLS; c += 32) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(b * TCOLS + c), v);
                if (c + 32 >= TCOLS) {      // last read of this accumulator: hand it back to the MMA issuer
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty(b));
                }
                if (BN >= 32 && n0 + c + 32 <= p.co) {
                    if (p.bias) {
#pragma unroll
                        for (int j = 0; j < 32; j++) v[j] += __ldg(p.bias + n0 + c + j);
                    }
#pragma unroll
                    for (int j = 0; j < 32; j += 4)
                        *reinterpret_cast<float4*>(stg + lane * STG + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
                    __syncwarp();
                    float4 t4[8];
#pragma unroll
                    for (int it = 0; it < 8; it++)
                        t4[it] = *reinterpret_cast<const float4*>(stg + (it * 4 + (lane >> 3)) * STG + (lane & 7) * 4);
#pragma unroll
                    for (int it = 0; it < 8; it++)
                        if ((rowok >> it) & 1u) *reinterpret_cast<float4*>(rowp[it] + c) = t4[it];
                    if (p.stats) {
                        float s1 = 0.f, s2 = 0.f;
#pragma unroll
                        for (int rr = 0; rr < 32; rr++) {
                            const float x = ((vmask >> rr) & 1u) ? stg[rr * STG + lane] : 0.f;
                            s1 += x; s2 = fmaf(x, x, s2);
                        }
                        red[(ew * TCOLS + c + lane) * 2 + 0] = s1;
                        red[(ew * TCOLS + c + lane) * 2 + 1] = s2;
                    }
                    __syncwarp();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; j++) {
                        const bool cv = n0 + c + j < p.co;
                        v[j] = cv ? v[j] + (p.bias ? __ldg(p.bias + n0 + c + j) : 0.f) : 0.f;
                        if (valid && cv) yrow[c + j] = v[j];
                    }
                    if (p.stats) {
                        float sq[32];
#pragma unroll
                        for (int j = 0; j < 32; j++) {
                            v[j] = valid ? v[j] : 0.f;
                            sq[j] = v[j] * v[j];
                        }
                        const float s1 = col_reduce32(v, lane);
                        const float s2 = col_reduce32(sq, lane);
                        red[(ew * TCOLS + c + lane) * 2 + 0] = s1;
                        red[(ew * TCOLS + c + lane) * 2 + 1] = s2;
                    }
                }
            }
            if (p.stats) {
                asm volatile("bar.sync 1, 128;" ::: "memory");      // the four epilogue warps only
                for (int col = threadIdx.x - 128; col < TCOLS; col += 128) {
                    if (n0 + col >= p.co) continue;
                    float s1 = 0.f, s2 = 0.f;
#pragma unroll
                    for (int w = 0; w < 4; w++) { s1 += red[(w * TCOLS + col) * 2]; s2 += red[(w * TCOLS + col) * 2 + 1]; }
                    double* dst = p.stats + ((long long)(p.stats_per_n ? n : 0) * p.co + n0 + col) * 2;
                    atomicAdd(dst, (double)s1);
                    atomicAdd(dst + 1, (double)s2);
                }
                asm volatile("bar.sync 1, 128;" ::: "memory");      // red is rewritten by the next unit
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<2 * TCOLS>(tmem_base);
    }
}

template <int BN>
static int launch_halo_persist(const AMaps& amaps, const CUtensorMap& w_hi, const CUtensorMap& w_lo, TcPersistP& pp, cudaStream_t st) {
    const int W_STAGE = (pp.h.terms == 2 ? 1 : 2) * BN * 128;
    constexpr int TCOLS = BN < 32 ? 32 : BN;
    constexpr int EPI_BYTES = 4 * 32 * 36 * 4 + 4 * TCOLS * 2 * 4;
    constexpr int MAX_SMEM = 227 * 1024;
    TcHaloP& p = pp.h;
    p.na = NA_MAX;
    const int fixed = p.na * 2 * p.a_plane + EPI_BYTES + 1024 + 512;
    int nw = (MAX_SMEM - fixed) / W_STAGE;
    if (nw > 8) nw = 8;
    if (nw < 2) return SKIT_ERR_UNSUPPORTED;
    p.nw = nw;
    const int smem = fixed + nw * W_STAGE;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_halo_persist_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(conv_tc_halo_persist_kernel<%d>) failed: %s", BN, cudaGetErrorString(e));
            return SKIT_ERR_CUDA;
        }
        attr_set = true;
    }
    const int grid = pp.units < 148 ? pp.units : 148;
    launch_pdl(conv_tc_halo_persist_kernel<BN>, grid, 256, smem, st, amaps, w_hi, w_lo, pp);
    return check_launch("conv_tc_halo_persist_kernel");
}


// ---------------------------------------------------------------------------------------------------------------------
// "Pixels on N" variant for layers with at most 128 output channels (the full-resolution 64 / 128-channel convs of the
// generator and their input gradients, the discriminators' stride-2 input gradients):
//
//   D^T[co][pixel] = sum_{chunk, tap, c} W[tap][co][chunk*64 + c] * A[pixel + tap][chunk*64 + c]
//
// The filter tile (128 output-channel rows, TMA zero-fills rows beyond co) is the UMMA A operand and the activation halo tile the
// B operand with N = 8 x TY pixels (TY = 16 / 24 / 32 rows of 8 pixels: the same row-granular descriptor trick as above, any
// number of 8-row groups).  One M = 128, N = 256 instruction does the work of two N = 128 (or four N = 64) instructions of the
// pixel-major kernels, whose cost per instruction barely falls with N (measured ~145 / 128 / 96 cycles for N = 256 / 128 / 64:
// below N = 256 the shared-memory operand fetch, not the MMA floor of max(M,128)*N/256 cycles, sets the pace).
// The accumulator comes out channel-major — TMEM lane = output channel, column = pixel — so the epilogue needs no staging
// at all: a warp's 32 lanes hold 32 consecutive channels of one pixel (one coalesced 128-byte store per pixel), the bias is a
// per-thread scalar and the norm statistics are per-thread running sums, flushed as fp64 atomics only when the CTA moves to
// another (image, channel tile).  Persistent: one CTA per SM, two 256-column accumulators, epilogue of unit i under the main loop
// of unit i + 1.  Warp roles (256 threads): 0 = filter TMA, 1 = MMA issuer + TMEM allocator, 2 = activation TMA, 4..7 = epilogue.
struct TcTransP {
    TcHaloP h;            // single region in h.reg[0]; h.a_plane holds the (8 + kw - 1) x (ty_rows + kh - 1) halo tile
    int ty_rows;          // output rows per tile: N = 8 * ty_rows
    int w_stage;          // bytes per filter stage: 2 planes of 128 rows x 128 B (three-term) or 1 (two-term)
    int tiles, mtiles, n_img, units;
};

__global__ void __launch_bounds__(256, 1)
conv_tc_halo_t_kernel(const __grid_constant__ AMaps tmA, const __grid_constant__ CUtensorMap tmW_hi,
                      const __grid_constant__ CUtensorMap tmW_lo, const TcTransP pp) {
    pdl_trigger();       // PDL: the next kernel in the stream may be scheduled once every CTA of this grid has started
    constexpr int BM = 128;
    constexpr int W_PLANE = BM * 128;
    constexpr int ACC_COLS = 256;
    const TcHaloP& p = pp.h;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
    const int a_stage = 2 * p.a_plane;
    const int NA = p.na;
    const uint32_t w0 = smem0 + NA * a_stage;
    const uint32_t bar0 = w0 + p.nw * pp.w_stage;
    auto a_full = [&](int s) { return bar0 + 8u * s; };
    auto a_empty = [&](int s) { return bar0 + 8u * (NA + s); };
    auto w_full = [&](int s) { return bar0 + 8u * (2 * NA + s); };
    auto w_empty = [&](int s) { return bar0 + 8u * (2 * NA + p.nw + s); };
    auto acc_full = [&](int b) { return bar0 + 8u * (2 * NA + 2 * p.nw + b); };
    auto acc_empty = [&](int b) { return bar0 + 8u * (2 * NA + 2 * p.nw + 2 + b); };
    const uint32_t tmem_slot = bar0 + 8u * (2 * NA + 2 * p.nw + 4);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + (tmem_slot - smem0));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const HaloRegion R = p.reg[0];
    const CUtensorMap* tmA_hi = &tmA.hi[0];
    const CUtensorMap* tmA_lo = &tmA.lo[0];
    const int TY = pp.ty_rows;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(tmA_hi); tma_prefetch_desc(tmA_lo);
        tma_prefetch_desc(&tmW_hi); tma_prefetch_desc(&tmW_lo);
        for (int s = 0; s < NA; s++) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
        for (int s = 0; s < p.nw; s++) { mbar_init(w_full(s), 1); mbar_init(w_empty(s), 1); }
        for (int b = 0; b < 2; b++) { mbar_init(acc_full(b), 1); mbar_init(acc_empty(b), 4); }
        mbar_fence_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc<2 * ACC_COLS>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();          // PDL: barrier init / TMEM allocation / descriptor prefetch above overlap the predecessor's tail
    const uint32_t tmem_base = *tmem_slot_gen;
    const int tiles_per_n = pp.tiles * pp.mtiles;

    // unit u -> (image, channel tile, pixel tile): pixel tiles fastest, so a CTA's consecutive units share (image, channel tile)
    auto decode = [&](int u, int& n, int& m0, int& y0, int& x0) {
        n = u / tiles_per_n;
        const int r = u - n * tiles_per_n;
        const int mt = r / pp.tiles, t = r - mt * pp.tiles;
        m0 = mt * BM;
        const int ty = t / R.tiles_x;
        y0 = ty * TY; x0 = (t - ty * R.tiles_x) * 8;
    };

    if (warp == 2 && lane == 0) {
        // ---------------- activation producer: one halo box (hi + lo) per 64-channel chunk
        const uint32_t bytes = 2u * (uint32_t)R.a_rows * 128u;
        int it = 0;
        for (int u = blockIdx.x; u < pp.units; u += gridDim.x) {
            int n, m0, y0, x0;
            decode(u, n, m0, y0, x0);
            for (int c = 0; c < p.kc; c++, it++) {
                const int s = it % NA, ph = (it / NA) & 1;
                mbar_wait(a_empty(s), ph ^ 1);
                mbar_expect_tx(a_full(s), bytes);
                const uint32_t sa = smem0 + s * a_stage;
                tma_load_4d(sa, tmA_hi, a_full(s), c * 64, R.org_x + x0, R.org_y + y0, n);
                tma_load_4d(sa + p.a_plane, tmA_lo, a_full(s), c * 64, R.org_x + x0, R.org_y + y0, n);
            }
        }
    } else if (warp == 0 && lane == 0) {
        // ---------------- filter producer: (chunk, tap) stages of 128 output-channel rows
        int it = 0;
        for (int u = blockIdx.x; u < pp.units; u += gridDim.x) {
            int n, m0, y0, x0;
            decode(u, n, m0, y0, x0);
            for (int c = 0; c < p.kc; c++)
                for (int ky = 0; ky < R.kh; ky++)
                    for (int kx = 0; kx < R.kw; kx++, it++) {
                        const int s = it % p.nw, ph = (it / p.nw) & 1;
                        const int tap = R.tap_base + ky * R.tap_sy + kx * R.tap_sx;
                        mbar_wait(w_empty(s), ph ^ 1);
                        mbar_expect_tx(w_full(s), p.terms == 2 ? W_PLANE : 2 * W_PLANE);
                        const uint32_t sw = w0 + s * pp.w_stage;
                        tma_load_3d(sw, &tmW_hi, w_full(s), c * 64, m0, tap);
                        if (p.terms != 2) tma_load_3d(sw + W_PLANE, &tmW_lo, w_full(s), c * 64, m0, tap);
                    }
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: D^T[co][pixel] += W[co][k] * A[pixel][k], accumulator (unit index & 1).  Whole warp,
        // warp-uniform control flow, one elected lane issues (see conv_tc_halo_kernel)
        const uint32_t idesc = make_idesc_bf16_mn(BM, 8 * TY, 0, 0);
        const uint32_t sbo_x = (uint32_t)R.pitch * 128u;
        const uint64_t dX = make_desc(0, 16, sbo_x), dW = make_desc(0, 16, 1024);
        const bool leader = elect_one();
        int ita = 0, itw = 0, ui = 0;
        for (int u = blockIdx.x; u < pp.units; u += gridDim.x, ui++) {
            const int b = ui & 1;
            mbar_wait(acc_empty(b), ((ui >> 1) & 1) ^ 1);
            tc_fence_after();
            const uint32_t dacc = tmem_base + (uint32_t)(b * ACC_COLS);
            uint32_t acc = 0;
            for (int c = 0; c < p.kc; c++, ita++) {
                const int s = ita % NA, ph = (ita / NA) & 1;
                const int kkc = (c == p.kc - 1) ? p.kk_last : 4;
                mbar_wait(a_full(s), ph);
                tc_fence_after();
                const uint32_t sa = smem0 + s * a_stage;
                for (int ky = 0; ky < R.kh; ky++)
                    for (int kx = 0; kx < R.kw; kx++, itw++) {
                        const uint32_t xrow = sa + (uint32_t)(ky * R.pitch + kx) * 128u;
                        const int ws = itw % p.nw, wph = (itw / p.nw) & 1;
                        mbar_wait(w_full(ws), wph);
                        tc_fence_after();
                        const uint32_t sw = w0 + ws * pp.w_stage;
                        if (leader) {
                            const uint64_t x_hi = dX + (xrow >> 4), x_lo = dX + ((xrow + (uint32_t)p.a_plane) >> 4);
                            const uint64_t w_hi = dW + (sw >> 4), w_lo = dW + ((sw + W_PLANE) >> 4);
                            if (p.terms == 2) {
#pragma unroll 4
                                for (int kk = 0; kk < kkc; kk++) {
                                    mma_bf16(dacc, w_hi + 2 * kk, x_lo + 2 * kk, idesc, acc);
                                    acc = 1u;
                                    mma_bf16(dacc, w_hi + 2 * kk, x_hi + 2 * kk, idesc, 1u);
                                }
                            } else {
#pragma unroll 4
                                for (int kk = 0; kk < kkc; kk++) {
                                    mma_bf16(dacc, w_hi + 2 * kk, x_lo + 2 * kk, idesc, acc);
                                    acc = 1u;
                                    mma_bf16(dacc, w_lo + 2 * kk, x_hi + 2 * kk, idesc, 1u);
                                    mma_bf16(dacc, w_hi + 2 * kk, x_hi + 2 * kk, idesc, 1u);
                                }
                            }
                            mma_commit(w_empty(ws));
                        }
                        __syncwarp();
                    }
                if (leader) mma_commit(a_empty(s));
                __syncwarp();
            }
            if (leader) mma_commit(acc_full(b));
            __syncwarp();
        }
    } else if (warp >= 4) {
        // ---------------- epilogue warps: TMEM lane = output channel m0 + 32*(warp-4) + lane, column = pixel ty*8 + tx of the tile
        const int ew = warp - 4;
        const int ncols = 8 * TY;
        double a1 = 0.0, a2 = 0.0;          // this thread's channel: running sum / sum of squares for the current (image, channel tile)
        int cur_n = -1, cur_m0 = -1;
        auto flush = [&]() {
            if (p.stats && cur_n >= 0) {
                const int ch = cur_m0 + ew * 32 + lane;
                if (ch < p.co) {
                    double* dst = p.stats + ((long long)(p.stats_per_n ? cur_n : 0) * p.co + ch) * 2;
                    atomicAdd(dst, a1);
                    atomicAdd(dst + 1, a2);
                }
            }
            a1 = 0.0; a2 = 0.0;
        };
        int ui = 0;
        for (int u = blockIdx.x; u < pp.units; u += gridDim.x, ui++) {
            int n, m0, y0, x0;
            decode(u, n, m0, y0, x0);
            if (n != cur_n || m0 != cur_m0) { flush(); cur_n = n; cur_m0 = m0; }
            const int b = ui & 1;
            const int ch = m0 + ew * 32 + lane;
            const bool chv = ch < p.co;
            const float bv = (p.bias && chv) ? __ldg(p.bias + ch) : 0.f;
            mbar_wait(acc_full(b), (ui >> 1) & 1);
            tc_fence_after();
            float s1 = 0.f, s2 = 0.f;
#pragma unroll 1
            for (int c = 0; c < ncols; c += 32) {
                float v[32];
                tmem_ld32(tmem_base + ((uint32_t)(ew * 32) << 16) + (uint32_t)(b * ACC_COLS + c), v);
                if (c + 32 >= ncols) {      // last read of this accumulator: hand it back to the MMA issuer
                    tc_fence_before();
                    __syncwarp();
                    if (lane == 0) mbar_arrive(acc_empty(b));
                }
#pragma unroll
                for (int r = 0; r < 4; r++) {           // 4 image rows of 8 pixels per 32-column chunk
                    const int oy = y0 + (c >> 3) + r;
                    const int py = (oy + R.ooy) * p.osy + p.poy;
                    const bool rv = chv && oy < R.ho && py < p.OH;
                    float* rowp = p.y + (((long long)n * p.OH + py) * p.OW + (long long)(x0 + R.oox) * p.osx + p.pox) * p.co + ch;
#pragma unroll
                    for (int tx = 0; tx < 8; tx++) {
                        const int ox = x0 + tx;
                        const int px = (ox + R.oox) * p.osx + p.pox;
                        if (rv && ox < R.wo && px < p.OW) {
                            const float val = v[r * 8 + tx] + bv;
                            rowp[(long long)tx * p.osx * p.co] = val;
                            s1 += val; s2 = fmaf(val, val, s2);
                        }
                    }
                }
            }
            a1 += (double)s1; a2 += (double)s2;
        }
        flush();
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<2 * ACC_COLS>(tmem_base);
    }
}

static bool trans_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SKIT_TC_TRANS");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

// Rows per tile and stage counts that fit 227 KB: prefer N = 256 (TY = 32) with two activation stages and >= 2 filter stages,
// then N = 192; a single activation stage only as the last resort.  Returns 0 when nothing fits.
static int pick_trans_shape(int pitch, int kh, int w_stage, int* a_plane, int* na, int* nw) {
    constexpr int MAX_SMEM = 227 * 1024;
    const int cand[2] = {32, 24};
    // tuning aid: SKIT_TRANS_TY / SKIT_TRANS_NA force a shape (tools/bench_trans.py sweeps them)
    const char* ety = getenv("SKIT_TRANS_TY");
    const char* ena = getenv("SKIT_TRANS_NA");
    if (ety && ena) {
        const int TY = atoi(ety), fna = atoi(ena);
        const int ap = ((pitch * (TY + kh - 1) * 128 + 1023) / 1024) * 1024;
        int n = (MAX_SMEM - (fna * 2 * ap + 2048)) / w_stage;
        if (n > 6) n = 6;
        if ((TY == 16 || TY == 24 || TY == 32) && (fna == 1 || fna == 2) && n >= 2) { *a_plane = ap; *na = fna; *nw = n; return TY; }
    }
    for (int i = 0; i < 2; i++)
        for (int want_na = 2; want_na >= 1; want_na--) {      // N = 256 first (measured never slower than N = 192), two activation stages if they fit
            const int TY = cand[i];
            const int ap = ((pitch * (TY + kh - 1) * 128 + 1023) / 1024) * 1024;
            const int fixed = want_na * 2 * ap + 1024 + 1024;
            int n = (MAX_SMEM - fixed) / w_stage;
            if (n > 6) n = 6;
            if (n >= (want_na == 2 ? 2 : 3)) { *a_plane = ap; *na = want_na; *nw = n; return TY; }
        }
    return 0;
}

static int launch_halo_t(const AMaps& amaps, const CUtensorMap& w_hi, const CUtensorMap& w_lo, TcTransP& pp, cudaStream_t st) {
    constexpr int MAX_SMEM = 227 * 1024;
    TcHaloP& p = pp.h;
    const int smem = p.na * 2 * p.a_plane + p.nw * pp.w_stage + 1024 + 1024;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_halo_t_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(conv_tc_halo_t_kernel) failed: %s", cudaGetErrorString(e));
            return SKIT_ERR_CUDA;
        }
        attr_set = true;
    }
    const int grid = pp.units < 148 ? pp.units : 148;
    launch_pdl(conv_tc_halo_t_kernel, grid, 256, smem, st, amaps, w_hi, w_lo, pp);
    return check_launch("conv_tc_halo_t_kernel");
}

long long* g_dbg_buffer = nullptr;

int encode_bf16_map_sw(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                       const uint32_t* box, CUtensorMapSwizzle swz) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return SKIT_ERR_CUDA;
    }
    cuuint64_t gd[5], gs[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; i++) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = 1; }
    for (int i = 0; i + 1 < rank; i++) gs[i] = strides_bytes[i];
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, swz, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=[%llu,%llu,%llu] box=[%u,%u,%u]", (int)r, rank,
                  (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
                  box[0], box[1], rank > 2 ? box[2] : 0);
        return SKIT_ERR_CUDA;
    }
    return SKIT_OK;
}

template <int BN>
static int launch_halo(const AMaps& amaps, const CUtensorMap& w_hi, const CUtensorMap& w_lo, TcHaloP& p, dim3 grid, cudaStream_t st) {
    const int W_STAGE = (p.terms == 2 ? 1 : 2) * BN * 128;
    constexpr int MAX_SMEM = 227 * 1024;
    // two activation stages unless a large halo tile would squeeze the weight ring below two stages (k4 with BN = 256)
    p.na = (MAX_SMEM - (NA_MAX * 2 * p.a_plane + 1024 + 512)) / W_STAGE >= 2 ? NA_MAX : 1;
    const int fixed = p.na * 2 * p.a_plane + 1024 + 512;
    int nw = (MAX_SMEM - fixed) / W_STAGE;
    if (nw > 8) nw = 8;
    if (nw < 2) {
        set_error("conv_tc_halo: the halo tile (%d bytes per plane) leaves no room for weight stages", p.a_plane);
        return SKIT_ERR_UNSUPPORTED;
    }
    p.nw = nw;
    const int smem = fixed + nw * W_STAGE;
    static int attr_smem = 0;
    if (smem > attr_smem) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_halo_kernel<BN>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(conv_tc_halo_kernel<%d>) failed: %s", BN, cudaGetErrorString(e));
            return SKIT_ERR_CUDA;
        }
        attr_smem = MAX_SMEM;
    }
    launch_pdl(conv_tc_halo_kernel<BN>, grid, 128, smem, st, amaps, w_hi, w_lo, p);
    return check_launch("conv_tc_halo_kernel");
}

// N tile for `tiles` pixel tiles of `ksteps` (chunk, tap) steps each: as wide as possible (the weight stream is the L2 traffic
// that remains), narrower when that fills more SMs.  Cost per CTA: K steps x max(MMA, shared-memory reads feeding them)
// + prologue and epilogue; total = waves x per-CTA cost.
static int pick_bn(int co, long long tiles, double ksteps, int a_plane) {
    if (co <= 16) return 16;
    int BN = 64;
    double best = 1e30;
    const int avail1 = 227 * 1024 - (2 * a_plane + 1024 + 512);   // with a single activation stage
    for (int cand = 256; cand >= 64; cand >>= 1) {
        if (cand > 64 && co % cand) continue;
        if (cand > 64 && avail1 / (cand * 256) < 2) continue;      // needs two weight stages next to the halo tile
        const int ntile = cdiv(co, cand);
        const double mma = cand * 0.53, rd = (4096.0 + cand * 32.0) / 80.0;
        const double per = ksteps * (12.0 * (mma > rd ? mma : rd) + 100.0) + 3000.0 + 16.0 * cand;
        const double cost = (double)((tiles * ntile + 147) / 148) * per;
        if (cost < best * 0.97) { best = cost; BN = cand; }
    }
    return BN;
}

static int encode_a_maps(tc::AMaps* am, int slot, const skit_operand* x, int pitch, int rows) {
    uint64_t dims[4] = {(uint64_t)x->c, (uint64_t)x->wp, (uint64_t)x->hp, (uint64_t)x->n};
    uint64_t strides[3] = {(uint64_t)x->c * 2, (uint64_t)x->c * 2 * x->wp, (uint64_t)x->c * 2 * x->wp * x->hp};
    uint32_t box[4] = {64, (uint32_t)pitch, (uint32_t)rows, 1};
    int rc = tc::encode_bf16_map_sw(&am->hi[slot], x->p0, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    return tc::encode_bf16_map_sw(&am->lo[slot], x->p1, 4, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

static int encode_w_maps(CUtensorMap* m_hi, CUtensorMap* m_lo, const void* w_hi, const void* w_lo, int ci_pack, int co, int ntaps_total, int BN) {
    uint64_t dims[3] = {(uint64_t)ci_pack, (uint64_t)co, (uint64_t)ntaps_total};
    uint64_t strides[2] = {(uint64_t)ci_pack * 2, (uint64_t)ci_pack * 2 * co};
    uint32_t box[3] = {64, (uint32_t)BN, 1};
    int rc = tc::encode_bf16_map_sw(m_hi, w_hi, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
    if (rc) return rc;
    // two-term launches (w_lo == NULL) never touch the lo map: give it a valid encoding of the hi plane
    return tc::encode_bf16_map_sw(m_lo, w_lo ? w_lo : w_hi, 3, dims, strides, box, CU_TENSOR_MAP_SWIZZLE_128B);
}

static int dbg_mode_env() {
    static int v = -1;
    if (v < 0) { const char* e = getenv("SKIT_DBG_MODE"); v = e ? atoi(e) : 0; }
    return v;
}

static bool persist_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SKIT_TC_PERSIST");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

static int dispatch_halo(int BN, const tc::AMaps& am, const CUtensorMap& m_hi, const CUtensorMap& m_lo, tc::TcHaloP& p, dim3 grid, cudaStream_t st) {
    using namespace tc;
    if (BN == 256) return launch_halo<256>(am, m_hi, m_lo, p, grid, st);
    if (BN == 128) return launch_halo<128>(am, m_hi, m_lo, p, grid, st);
    if (BN == 64) return launch_halo<64>(am, m_hi, m_lo, p, grid, st);
    return launch_halo<16>(am, m_hi, m_lo, p, grid, st);
}

}  // namespace tc

// Stride-1 valid convolution of a haloed bf16x2 operand.  kh x kw taps of THIS launch; ntaps_total = taps in the
// packed filter [tap][co][ci_pack]; ci_pack = channel count of the weight pack rows (>= x->c, multiple of 8).
int conv_tc_halo_launch(const skit_operand* x, const void* w_hi, const void* w_lo, int ci_pack, int co, int kh, int kw,
                        int ntaps_total, int tap_base, int org, int ho, int wo, const float* bias, float* y,
                        const TcOut* out, double* stats, int stats_mode, cudaStream_t st) {
    using namespace tc;
    const int ci = x->c;
    TcHaloP p{};
    p.kc = cdiv(ci, 64);
    p.kk_last = cdiv(ci - (p.kc - 1) * 64, 16);
    p.co = co;
    if (out) { p.OH = out->OH; p.OW = out->OW; p.osy = out->osy; p.osx = out->osx; p.poy = out->ooy; p.pox = out->oox; }
    else { p.OH = ho; p.OW = wo; p.osy = 1; p.osx = 1; p.poy = 0; p.pox = 0; }
    p.bias = bias; p.y = y; p.stats = stats; p.stats_per_n = stats_mode == SKIT_NORM_INSTANCE;
    p.dbg = g_dbg_buffer;
    p.dbg_mode = dbg_mode_env();
    p.terms = w_lo ? 3 : 2;
    p.nreg = 1;
    HaloRegion& R = p.reg[0];
    R.first = 0; R.kh = kh; R.kw = kw; R.tap_base = tap_base; R.tap_sy = kw; R.tap_sx = 1;
    R.org_y = org; R.org_x = org; R.ho = ho; R.wo = wo; R.tiles_x = cdiv(wo, 8); R.ooy = 0; R.oox = 0; R.amap = 0;
    R.pitch = 8 + kw - 1; R.a_rows = R.pitch * (16 + kh - 1);
    p.a_plane = ((R.a_rows * 128 + 1023) / 1024) * 1024;
    const long long pm_tiles = (long long)R.tiles_x * cdiv(ho, 16) * x->n;      // tiles of the pixel-major kernels
    static long long big_min = -1;      // pixel-major tiles from which 256 / 512-channel layers take the pixels-on-N kernel (SKIT_TRANS_MIN_PM_TILES)
    if (big_min < 0) { const char* e = getenv("SKIT_TRANS_MIN_PM_TILES"); big_min = e ? atoll(e) : 4 * 148; }
    if (trans_enabled() && co >= 32 && (co <= 128 || (co <= 512 && co % 128 == 0 && pm_tiles >= big_min)) && !p.dbg) {
        // channels on M (tiles of 128), 192 / 256 pixels on N (conv_tc_halo_t_kernel) when the map fills the SMs.  For 256 / 512
        // output channels it is the same MMA work as the pixel-major N = 256 kernel with half the filter streaming per pixel (a
        // unit streams only its 128 filter rows); measured (tools/bench_trans.py, 768x768 step): 128 -> 256 at 384x384 208 -> 194 us,
        // its input gradient 247 -> 170 us, but the 256 -> 256 trunk at 192x192 (two waves of pixel-major tiles) 88 -> 94 us: only
        // maps of four waves or more take this path above 128 channels
        TcTransP pp{};
        const int w_stage = (p.terms == 2 ? 1 : 2) * 128 * 128;
        int ap = 0, na = 0, nw = 0;
        const int TY = pick_trans_shape(R.pitch, kh, w_stage, &ap, &na, &nw);
        const long long tiles_t = TY ? (long long)R.tiles_x * cdiv(ho, TY) : 0;
        const int mtiles = cdiv(co, 128);
        if (TY && tiles_t * x->n * mtiles >= 120 && tiles_t * x->n * mtiles < (1ll << 30)) {
            pp.h = p;
            pp.h.a_plane = ap; pp.h.na = na; pp.h.nw = nw;
            pp.h.reg[0].a_rows = R.pitch * (TY + kh - 1);
            pp.ty_rows = TY; pp.w_stage = w_stage;
            pp.tiles = (int)tiles_t; pp.mtiles = mtiles; pp.n_img = x->n; pp.units = (int)(tiles_t * x->n * mtiles);
            AMaps amt;
            CUtensorMap t_hi, t_lo;
            int rc = encode_a_maps(&amt, 0, x, R.pitch, TY + kh - 1);
            if (rc) return rc;
            for (int i = 1; i < MAX_AMAPS; i++) { amt.hi[i] = amt.hi[0]; amt.lo[i] = amt.lo[0]; }
            rc = encode_w_maps(&t_hi, &t_lo, w_hi, w_lo, ci_pack, co, ntaps_total, 128);
            if (rc) return rc;
            return launch_halo_t(amt, t_hi, t_lo, pp, st);
        }
    }
    const int tiles_y = cdiv(ho, 16);
    const int BN = pick_bn(co, (long long)R.tiles_x * tiles_y * x->n, (double)p.kc * kh * kw, p.a_plane);
    AMaps am;
    CUtensorMap m_hi, m_lo;
    int rc = encode_a_maps(&am, 0, x, R.pitch, 16 + kh - 1);
    if (rc) return rc;
    for (int i = 1; i < MAX_AMAPS; i++) { am.hi[i] = am.hi[0]; am.lo[i] = am.lo[0]; }
    rc = encode_w_maps(&m_hi, &m_lo, w_hi, w_lo, ci_pack, co, ntaps_total, BN);
    if (rc) return rc;
    dim3 grid(R.tiles_x * tiles_y, cdiv(co, BN), x->n);
    const long long units = (long long)grid.x * grid.y * grid.z;
    if (BN <= 128 && units >= 2 * 148 && units < (1ll << 30) && persist_enabled() && !p.dbg) {
        // several waves of short tiles: persistent CTAs with two accumulators overlap each tile's epilogue with the next main loop
        TcPersistP pp{};
        pp.h = p;
        pp.tiles = grid.x; pp.ntiles = grid.y; pp.n_img = grid.z; pp.units = (int)units;
        int prc = SKIT_ERR_UNSUPPORTED;
        if (BN == 128) prc = launch_halo_persist<128>(am, m_hi, m_lo, pp, st);
        else if (BN == 64) prc = launch_halo_persist<64>(am, m_hi, m_lo, pp, st);
        else prc = launch_halo_persist<16>(am, m_hi, m_lo, pp, st);
        if (prc != SKIT_ERR_UNSUPPORTED) return prc;
    }
    return dispatch_halo(BN, am, m_hi, m_lo, p, grid, st);
}

// Stride-1 INPUT GRADIENT of a k x k conv: dx[n][H][W][co] = full correlation of the gradient operand d (zero halo of k-1, so
// d->hp = H + k - 1 ... i.e. dx is (d->hp - k + 1) x (d->wp - k + 1)) with the flipped filter pack [k*k][co][ci].  When the
// one-region tiling would spill into an extra wave and the interior + four border strips fit better, it is launched as five
// regions of ONE grid: the strips' windows reach the data only through one filter row / column (the rest of the window is the
// zero halo), so they run a third of the K loop.
int conv_tc_halo_dgrad_full(const skit_operand* d, const void* w_hi, const void* w_lo, int co, int k, float* dx, cudaStream_t st) {
    using namespace tc;
    const int ci = d->c;
    const int H = d->hp - k + 1, W = d->wp - k + 1;
    const int q = k - 1;
    const long long full_tiles = (long long)cdiv(W, 8) * cdiv(H, 16) * d->n;
    const int hi_ = H - 2, wi_ = W - 2;          // interior: outputs whose window holds at least ... every tap row/col may matter
    const long long int_tiles = (long long)cdiv(wi_, 8) * cdiv(hi_, 16) * d->n;
    const long long strip_tiles = (long long)(2 * cdiv(W, 8) + 2 * cdiv(hi_, 16)) * d->n;
    // Cost in waves of 148 CTAs.  One region: ceil(full / 148).  Five regions: the interior's waves, plus — for the strips that do
    // not fit into the slack of its last wave — a third of a wave per 148 of them (a strip runs one filter row / column: a third
    // of the K loop).  Large maps (more than four waves) stay on the single-region path: it runs on the PERSISTENT kernel there
    // (one CTA per SM, epilogues overlapped), which beats thousands of one-tile CTAs paying their set-up each.
    const long long full_waves = (full_tiles + 147) / 148, int_waves = (int_tiles + 147) / 148;
    const long long slack = int_waves * 148 - int_tiles;
    const double split_cost = (double)int_waves + (strip_tiles > slack ? (double)((strip_tiles - slack + 147) / 148) / 3.0 : 0.0);
    static int split_on = -1;      // SKIT_DGRAD_SPLIT=0: always the single-region launch (-> pixels-on-N kernel), for A/B timing
    if (split_on < 0) { const char* e = getenv("SKIT_DGRAD_SPLIT"); split_on = (e && e[0] == '0') ? 0 : 1; }
    const bool split = split_on && k == 3 && d->n == 1 && H > 18 && W > 10 && ci % 64 == 0 && full_tiles <= 4 * 148 &&
                       split_cost < (double)full_waves - 0.05;
    if (!split)
        return conv_tc_halo_launch(d, w_hi, w_lo, ci, co, k, k, k * k, 0, 0, H, W, nullptr, dx, nullptr, nullptr, SKIT_NORM_NONE, st);
    TcHaloP p{};
    p.kc = cdiv(ci, 64); p.kk_last = 4; p.co = co;
    p.OH = H; p.OW = W; p.osy = 1; p.osx = 1; p.poy = 0; p.pox = 0;
    p.bias = nullptr; p.y = dx; p.stats = nullptr; p.stats_per_n = 0; p.dbg = g_dbg_buffer;
    p.dbg_mode = dbg_mode_env();
    p.terms = w_lo ? 3 : 2;
    p.nreg = 5;
    auto set = [&](int i, int kh, int kw, int tb, int sy, int sx, int oy, int ox, int ho, int wo, int ooy, int oox, int amap) {
        HaloRegion& R = p.reg[i];
        R.kh = kh; R.kw = kw; R.tap_base = tb; R.tap_sy = sy; R.tap_sx = sx; R.org_y = oy; R.org_x = ox; R.ho = ho; R.wo = wo;
        R.tiles_x = cdiv(wo, 8); R.ooy = ooy; R.oox = oox; R.amap = amap; R.pitch = 8 + kw - 1; R.a_rows = R.pitch * (16 + kh - 1);
    };
    // dx[i][x] = sum_{a,b} dz[i+a][x+b] * wf[a*3+b], dz = d with its zero halo of 2.  Row 0 sees data only through a = 2,
    // row H-1 only through a = 0; column 0 only through b = 2, column W-1 only through b = 0.
    set(0, 3, 3, 0, 3, 1, 1, 1, hi_, wi_, 1, 1, 0);                 // interior rows 1..H-2, cols 1..W-2, all 9 taps
    set(1, 1, 3, 6, 3, 1, 2, 0, 1, W, 0, 0, 1);                     // top row:    taps (2, b), operand row 0 + 2
    set(2, 1, 3, 0, 3, 1, H - 1, 0, 1, W, H - 1, 0, 1);             // bottom row: taps (0, b), operand row H-1 + 0
    set(3, 3, 1, 2, 3, 1, 1, 2, hi_, 1, 1, 0, 2);                   // left col:   taps (a, 2), operand col 0 + 2, rows 1..H-2
    set(4, 3, 1, 0, 3, 1, 1, W - 1, hi_, 1, 1, W - 1, 2);           // right col:  taps (a, 0)
    int first = 0, amax = 0;
    for (int i = 0; i < 5; i++) {
        p.reg[i].first = first;
        first += p.reg[i].tiles_x * cdiv(p.reg[i].ho, 16);
        amax = max(amax, p.reg[i].a_rows);
    }
    p.a_plane = ((amax * 128 + 1023) / 1024) * 1024;
    (void)q;
    const int BN = (co % 256 == 0) ? 256 : pick_bn(co, int_tiles, (double)p.kc * 9, p.a_plane);
    AMaps am;
    CUtensorMap m_hi, m_lo;
    int rc = encode_a_maps(&am, 0, d, 10, 18);
    if (rc) return rc;
    rc = encode_a_maps(&am, 1, d, 10, 16);
    if (rc) return rc;
    rc = encode_a_maps(&am, 2, d, 8, 18);
    if (rc) return rc;
    rc = encode_w_maps(&m_hi, &m_lo, w_hi, w_lo, ci, co, k * k, BN);
    if (rc) return rc;
    dim3 grid(first, cdiv(co, BN), 1);
    return dispatch_halo(BN, am, m_hi, m_lo, p, grid, st);
}

}  // namespace skit

/* Debug aid (not part of the product path): when set, the halo conv kernel writes clock64() stamps per CTA into
 * buf[cta][8] = {start, setup done, first A tile landed, last MMA issued, accumulator ready, epilogue stores done, end}. */
extern "C" int skit_debug_set_buffer(long long* buf) {
    skit::tc::g_dbg_buffer = buf;
    return SKIT_OK;
}
