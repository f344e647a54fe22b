// tcgen05 + TMA weight gradient:  dW[tap][m][n] = sum_{image, pixel} Mop[pixel (+tap)][m] * Nop[pixel (+tap)][n]
//
// Both GEMM operands are activations in NHWC (pixels are the reduction axis), i.e. MN-major in
// tcgen05 terms: a TMA box of 64 pixels x 64 channels lands as 64 rows of 128 B (SWIZZLE_128B) and
// is consumed directly — no transpose anywhere.  The M side is whichever of {output-gradient
// channels, input channels} is a multiple of 128 (one 128-row tile per CTA); the other side is the N
// tile (64/128/256).  The input-side box is shifted by the filter tap (and strided for stride-2
// layers through the tensor map's element strides); pixels outside the image read the zero halo of
// the gradient operand (or TMA zero fill), so ragged tiles need no masking.  fp32 fidelity: the same
// 3-term bf16 hi/lo split as the forward kernel.  Split-K over pixel tiles across CTAs; each CTA
// accumulates in TMEM and adds its partial to the fp32 result with coalesced red.global.add.
#include "tc_common.cuh"

namespace skit {
int bwd_terms();   // tc_conv.cu
namespace tc {

struct TcWgradP {
    int k;                 // filter rows
    int kw;                // filter columns (== k for square filters; 1 for x-folded operands)
    int tiles_x, tiles_y;  // 8x8-pixel tiles per image
    int n_img;
    int tiles_per_cta;     // split-K chunk (in pixel tiles, over all images)
    int total_tiles;
    int m_is_x;            // 1: M operand is the (tap-shifted, strided) input, N operand the gradient; 0: the reverse
    int x_org, x_stride, d_org, d_org_y;   // d_org: gradient-side column origin; d_org_y: its row origin
    int Mdim, Ndim;        // channel counts of the M / N side
    float* out;            // [tap][Ndim][Mdim] fp32 partial sums (zeroed by the caller)
    int skip;              // two-term product: 2 = the N operand's lo plane is neither loaded, staged nor multiplied; 0 = all three terms
    int stages;            // pipeline depth: what fits 227 KB (a two-term stage is N_BYTES smaller, so one more stage is in flight)
};

constexpr int PIX = 64;               // pixels per K stage (8x8 box)
constexpr int BOX_BYTES = PIX * 128;  // 64 pixels x 64 bf16

template <int BN, int STAGES_UNUSED>
__global__ void __launch_bounds__(128, 1)
wgrad_tc_kernel(const __grid_constant__ CUtensorMap tmM_hi, const __grid_constant__ CUtensorMap tmM_lo,
                const __grid_constant__ CUtensorMap tmN_hi, const __grid_constant__ CUtensorMap tmN_lo, TcWgradP p) {
    pdl_trigger();       // PDL: the next kernel in the stream may be scheduled once every CTA of this grid has started
    constexpr int M_BYTES = 2 * BOX_BYTES;          // 128 M-channels = 2 boxes (the second is TMA zero fill when Mdim == 64)
    constexpr int NBOX = BN >= 64 ? BN / 64 : 1;    // a thin N side (<= 16 channels) still lands as one 64-channel box, zero-filled
    constexpr int N_BYTES = NBOX * BOX_BYTES;
    constexpr int TCOLS = BN < 32 ? 32 : BN;
    const int STAGES = p.stages;
    const int STAGE_BYTES = 2 * M_BYTES + (p.skip == 2 ? 1 : 2) * N_BYTES;      // [M hi][M lo][N hi]([N lo])
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
    const uint32_t bar0 = smem0 + STAGES * STAGE_BYTES;
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    const uint32_t tmem_full_bar = bar0 + 8u * (2 * STAGES);
    const uint32_t tmem_slot = bar0 + 8u * (2 * STAGES + 1);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 1));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tap = blockIdx.x % (p.k * p.kw);
    const int mt = blockIdx.x / (p.k * p.kw);         // 128-row M tile
    const int n0 = blockIdx.y * BN;
    const int t_beg = blockIdx.z * p.tiles_per_cta;
    const int t_end = min(p.total_tiles, t_beg + p.tiles_per_cta);
    const int ky = tap / p.kw, kx = tap - ky * p.kw;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmM_hi); tma_prefetch_desc(&tmM_lo);
        tma_prefetch_desc(&tmN_hi); tma_prefetch_desc(&tmN_lo);
        for (int s = 0; s < STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
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
    const int num_it = t_end - t_beg;

    if (num_it > 0) {
        if (warp == 0 && lane == 0) {
            // ---------------- TMA producer
            const int tpi = p.tiles_x * p.tiles_y;
            for (int it = 0; it < num_it; it++) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(empty_bar(s), ph ^ 1);
                mbar_expect_tx(full_bar(s), STAGE_BYTES);
                const int t = t_beg + it;
                const int img = t / tpi, r = t - img * tpi;
                const int ty = r / p.tiles_x, tx = r - ty * p.tiles_x;
                const int xx = p.x_org + tx * 8 * p.x_stride + kx, xy = p.x_org + ty * 8 * p.x_stride + ky;  // input side
                const int dx = p.d_org + tx * 8, dy = p.d_org_y + ty * 8;                                       // gradient side
                const int mx = p.m_is_x ? xx : dx, my = p.m_is_x ? xy : dy;
                const int nx = p.m_is_x ? dx : xx, ny = p.m_is_x ? dy : xy;
                const uint32_t sa = smem0 + s * STAGE_BYTES;
#pragma unroll
                for (int g = 0; g < 2; g++) {
                    tma_load_4d(sa + g * BOX_BYTES, &tmM_hi, full_bar(s), mt * 128 + g * 64, mx, my, img);
                    tma_load_4d(sa + M_BYTES + g * BOX_BYTES, &tmM_lo, full_bar(s), mt * 128 + g * 64, mx, my, img);
                }
#pragma unroll
                for (int g = 0; g < NBOX; g++) {
                    tma_load_4d(sa + 2 * M_BYTES + g * BOX_BYTES, &tmN_hi, full_bar(s), n0 + g * 64, nx, ny, img);
                    if (p.skip != 2) tma_load_4d(sa + 2 * M_BYTES + N_BYTES + g * BOX_BYTES, &tmN_lo, full_bar(s), n0 + g * 64, nx, ny, img);
                }
            }
        } else if (warp == 1) {
            // ---------------- MMA issuer (both operands MN-major: LBO = one 64-channel box, SBO = 8 pixel rows).  Whole warp with
            // warp-uniform control flow, one elected lane issues (see conv_tc_halo_kernel); a K step of 16 pixels advances the
            // descriptors' start address by 16 rows x 128 B
            constexpr uint32_t idesc = make_idesc_bf16(BN, 1, 1);
            const uint64_t dsc = make_desc(0, BOX_BYTES, 1024);
            const bool leader = elect_one();
            for (int it = 0; it < num_it; it++) {
                const int s = it % STAGES, ph = (it / STAGES) & 1;
                mbar_wait(full_bar(s), ph);
                tc_fence_after();
                const uint32_t sa = smem0 + s * STAGE_BYTES;
                if (leader) {
                    const uint64_t m_hi = dsc + (sa >> 4), m_lo = dsc + ((sa + M_BYTES) >> 4);
                    const uint64_t n_hi = dsc + ((sa + 2 * M_BYTES) >> 4), n_lo = dsc + ((sa + 2 * M_BYTES + N_BYTES) >> 4);
                    constexpr uint64_t KS = (16 * 128) >> 4;
                    if (p.skip == 2) {
#pragma unroll
                        for (int kk = 0; kk < PIX / 16; kk++) {
                            mma_bf16(tmem_base, m_hi + KS * kk, n_hi + KS * kk, idesc, (it > 0 || kk > 0) ? 1u : 0u);
                            mma_bf16(tmem_base, m_lo + KS * kk, n_hi + KS * kk, idesc, 1u);
                        }
                    } else {
#pragma unroll
                        for (int kk = 0; kk < PIX / 16; kk++) {
                            mma_bf16(tmem_base, m_hi + KS * kk, n_hi + KS * kk, idesc, (it > 0 || kk > 0) ? 1u : 0u);
                            mma_bf16(tmem_base, m_lo + KS * kk, n_hi + KS * kk, idesc, 1u);
                            mma_bf16(tmem_base, m_hi + KS * kk, n_lo + KS * kk, idesc, 1u);
                        }
                    }
                    mma_commit(empty_bar(s));  // frees the smem stage when these MMAs retire
                }
                __syncwarp();
            }
            if (leader) mma_commit(tmem_full_bar);
        }
        __syncwarp();
        // ---------------- epilogue: out[(tap*Ndim + n)*Mdim + m] += acc   (lanes = consecutive m: coalesced reds)
        mbar_wait(tmem_full_bar, 0);
        tc_fence_after();
        __syncwarp();
        const int m = mt * 128 + warp * 32 + lane;
        float* obase = p.out + ((long long)tap * p.Ndim + n0) * p.Mdim + m;
#pragma unroll 1
        for (int c = 0; c < TCOLS; c += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
            if (m < p.Mdim) {
#pragma unroll
                for (int j = 0; j < 32; j++)
                    if (n0 + c + j < p.Ndim) atomicAdd(obase + (long long)(c + j) * p.Mdim, v[j]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<TCOLS>(tmem_base);
    }
}

template <int BN, int STAGES>
static int launch_wgrad_tc(const CUtensorMap& m_hi, const CUtensorMap& m_lo, const CUtensorMap& n_hi, const CUtensorMap& n_lo,
                           TcWgradP& p, dim3 grid, cudaStream_t st) {
    constexpr int MAX_SMEM = 227 * 1024;
    const int stage_bytes = 4 * BOX_BYTES + (p.skip == 2 ? 1 : 2) * (BN >= 64 ? BN / 64 : 1) * BOX_BYTES;
    int stages = (MAX_SMEM - 1024 - 256) / stage_bytes;
    if (stages > 6) stages = 6;
    if (stages < STAGES) stages = STAGES;      // the three-term depth always fits
    p.stages = stages;
    const int SMEM = stages * stage_bytes + 1024 + 256;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(wgrad_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, MAX_SMEM);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(wgrad_tc_kernel<%d>) failed: %s", BN, cudaGetErrorString(e));
            return SKIT_ERR_CUDA;
        }
        attr_set = true;
    }
    launch_pdl(wgrad_tc_kernel<BN, STAGES>, grid, 128, SMEM, st, m_hi, m_lo, n_hi, n_lo, p);
    return check_launch("wgrad_tc_kernel");
}

static int encode_act_map(CUtensorMap* hi, CUtensorMap* lo, const skit_operand* t, int estride) {
    uint64_t dims[4] = {(uint64_t)t->c, (uint64_t)t->wp, (uint64_t)t->hp, (uint64_t)t->n};
    uint64_t strides[3] = {(uint64_t)t->c * 2, (uint64_t)t->c * 2 * t->wp, (uint64_t)t->c * 2 * t->wp * t->hp};
    uint32_t box[4] = {64, (uint32_t)(8 * estride), (uint32_t)(8 * estride), 1};
    uint32_t es[4] = {1, (uint32_t)estride, (uint32_t)estride, 1};
    int rc = encode_bf16_map(hi, t->p0, 4, dims, strides, box, es);
    if (rc) return rc;
    return encode_bf16_map(lo, t->p1, 4, dims, strides, box, es);
}

}  // namespace tc

// layout of the partial-sum buffer the kernel fills: 0 = [tap][ci][co] (M = co), 1 = [tap][co][ci] (M = ci)
bool wgrad_tc_eligible(const skit_operand* x, const skit_operand* dy, int k, int stride, int ho, int wo) {
    if (x->fmt != SKIT_FMT_BF16X2 || dy->fmt != SKIT_FMT_BF16X2 || (stride != 1 && stride != 2)) return false;
    // M side: the larger channel count, a multiple of 64 (a 64-channel M side pads to the 128-row UMMA tile with TMA
    // zero fill); N side: a multiple of 64, or thin (<= 16 channels, a multiple of 8: N = 16 MMAs on a zero-filled box)
    const int ci = x->c, co = dy->c;
    const int big = ci > co ? ci : co, small = ci > co ? co : ci;
    if (big % 64) return false;
    return small % 64 == 0 || (small <= 16 && small % 8 == 0);
}

int wgrad_tc(const skit_operand* x, int org, const skit_operand* dy, int dy_org, int k, int stride,
             int ho, int wo, float* partial, int* layout, cudaStream_t st, int kw = 0, int dy_org_y = -1) {
    using namespace tc;
    const int ci = x->c, co = dy->c;
    TcWgradP p{};
    if (kw <= 0) kw = k;
    p.k = k; p.kw = kw;
    p.tiles_x = cdiv(wo, 8); p.tiles_y = cdiv(ho, 8);
    p.n_img = x->n;
    p.total_tiles = p.tiles_x * p.tiles_y * x->n;
    // M = whichever side fills 128-row tiles best: prefer a multiple of 128, else the larger one
    if (co % 128 == 0 && co >= ci) p.m_is_x = 0;
    else if (ci % 128 == 0 && ci >= co) p.m_is_x = 1;
    else if (co % 128 == 0 && ci % 64 == 0) p.m_is_x = 0;
    else if (ci % 128 == 0 && co % 64 == 0) p.m_is_x = 1;
    else p.m_is_x = ci > co ? 1 : 0;
    p.x_org = org; p.x_stride = stride; p.d_org = dy_org; p.d_org_y = dy_org_y >= 0 ? dy_org_y : dy_org;
    p.Mdim = p.m_is_x ? ci : co;
    p.Ndim = p.m_is_x ? co : ci;
    p.out = partial;
    p.skip = bwd_terms() == 2 ? 2 : 0;     // two-term: the N-side operand (the wider box set) enters as its hi plane only
    *layout = p.m_is_x;
    const int BN = (p.Ndim % 256 == 0) ? 256 : (p.Ndim % 128 == 0) ? 128 : (p.Ndim % 64 == 0) ? 64 : 16;
    const int mtiles = cdiv(p.Mdim, 128), ntiles = cdiv(p.Ndim, BN);
    const int items = k * kw * mtiles * ntiles;
    // split-K so that one wave of CTAs covers the SMs: every extra split adds a full tile of fp32 atomics to L2
    int splits = max(1, min(items >= 148 ? 1 : 148 / items, p.total_tiles));
    p.tiles_per_cta = cdiv(p.total_tiles, splits);
    splits = cdiv(p.total_tiles, p.tiles_per_cta);

    CUtensorMap x_hi, x_lo, d_hi, d_lo;
    int rc = encode_act_map(&x_hi, &x_lo, x, stride);
    if (rc) return rc;
    rc = encode_act_map(&d_hi, &d_lo, dy, 1);
    if (rc) return rc;
    dim3 grid(k * kw * mtiles, ntiles, splits);
    const CUtensorMap &mh = p.m_is_x ? x_hi : d_hi, &ml = p.m_is_x ? x_lo : d_lo;
    const CUtensorMap &nh = p.m_is_x ? d_hi : x_hi, &nl = p.m_is_x ? d_lo : x_lo;
    if (BN == 256) return launch_wgrad_tc<256, 2>(mh, ml, nh, nl, p, grid, st);
    if (BN == 128) return launch_wgrad_tc<128, 3>(mh, ml, nh, nl, p, grid, st);
    if (BN == 64) return launch_wgrad_tc<64, 4>(mh, ml, nh, nl, p, grid, st);
    return launch_wgrad_tc<16, 4>(mh, ml, nh, nl, p, grid, st);
}

}  // namespace skit
