// tcgen05 + TMA implicit-GEMM convolution (stride 1, valid, NHWC haloed bf16x2 operand).
//
//   D[pixel][co] = sum_{tap, ci} A[pixel + tap][ci] * W[tap][co][ci]
//
// GEMM view per CTA: M = 128 output pixels (a TH x TW patch of one image), N = BN output channels,
// K = taps * Ci walked in 64-channel steps.  Each K step is one TMA box of the haloed activation
// tensor (shifted by the tap: implicit im2col, nothing is materialised) and one box of the packed
// filter, both landing in 128B-swizzled shared memory.  fp32 fidelity comes from a 3-term bf16
// split: x = hi + lo for both operands and D += A_lo*W_hi + A_hi*W_lo + A_hi*W_hi, accumulated in
// fp32 in TMEM (a single bf16 pass misses the 1e-3 parity gate by 17x).
// Epilogue: TMEM -> registers -> (+bias) -> NHWC fp32 store, plus the per-channel sum / sum of
// squares of the tile (the InstanceNorm / BatchNorm statistics) reduced by warp shuffles and added
// to the global double accumulators.
//
// Warp roles (128 threads): warp 0 lane 0 = TMA producer, warp 1 lane 0 = MMA issuer,
// warp 1 = TMEM allocator, all four warps = epilogue (warp w owns TMEM lanes 32w..32w+31).
#include <cstdlib>
#include "tc_common.cuh"

namespace skit {
namespace tc {

EncodeTiledFn get_encode_tiled() {
    static EncodeTiledFn fn = nullptr;
    if (!fn) {
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = reinterpret_cast<EncodeTiledFn>(p);
    }
    return fn;
}

int encode_bf16_map(CUtensorMap* m, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                    const uint32_t* box, const uint32_t* estrides) {
    EncodeTiledFn enc = get_encode_tiled();
    if (!enc) {
        set_error("cuTensorMapEncodeTiled entry point not available");
        return SKIT_ERR_CUDA;
    }
    cuuint64_t gd[5];
    cuuint64_t gs[5];
    cuuint32_t bx[5], es[5];
    for (int i = 0; i < rank; i++) { gd[i] = dims[i]; bx[i] = box[i]; es[i] = estrides ? estrides[i] : 1; }
    for (int i = 0; i + 1 < rank; i++) gs[i] = strides_bytes[i];
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
        set_error("cuTensorMapEncodeTiled failed (CUresult %d) rank=%d dims=[%llu,%llu,%llu,%llu] box=[%u,%u,%u,%u]", (int)r,
                  rank, (unsigned long long)dims[0], (unsigned long long)dims[1], (unsigned long long)(rank > 2 ? dims[2] : 0),
                  (unsigned long long)(rank > 3 ? dims[3] : 0), box[0], box[1], rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
        return SKIT_ERR_CUDA;
    }
    return SKIT_OK;
}

struct TcConvP {
    int k, kc;        // filter size, Ci/64
    int org;          // halo origin offset
    int stride;       // input stride (TMA elementStrides do the striding)
    int tap_base;     // first tap of this launch inside the packed filter (phase packs)
    int ho, wo, co;
    int tw, th;       // pixel patch (tw*th == 128)
    int tiles_x;
    // output placement: y[n][oy*osy + ooy][ox*osx + oox][co] inside an [OH][OW] map (phase-wise dgrad writes)
    int OH, OW, osy, osx, ooy, oox;
    const float* bias;
    float* y;
    double* stats;
    int stats_per_n;
};

constexpr int A_BYTES = 128 * 128;  // 128 pixels x 64 bf16

template <int BN, int STAGES>
__global__ void __launch_bounds__(128, 1)
conv_tc_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
               const __grid_constant__ CUtensorMap tmW_hi, const __grid_constant__ CUtensorMap tmW_lo, TcConvP p) {
    pdl_trigger();       // PDL: the next kernel in the stream may be scheduled once every CTA of this grid has started
    constexpr int W_BYTES = BN * 128;
    constexpr int STAGE_BYTES = 2 * A_BYTES + 2 * W_BYTES;
    extern __shared__ uint8_t smem_raw[];
    const uint32_t smem0 = (smem_u32(smem_raw) + 1023u) & ~1023u;
    uint8_t* smem_gen = smem_raw + (smem0 - smem_u32(smem_raw));
    const uint32_t bar0 = smem0 + STAGES * STAGE_BYTES;  // full[STAGES], empty[STAGES], tmem_full, slot
    auto full_bar = [&](int s) { return bar0 + 8u * s; };
    auto empty_bar = [&](int s) { return bar0 + 8u * (STAGES + s); };
    const uint32_t tmem_full_bar = bar0 + 8u * (2 * STAGES);
    const uint32_t tmem_slot = bar0 + 8u * (2 * STAGES + 1);
    volatile uint32_t* tmem_slot_gen = reinterpret_cast<volatile uint32_t*>(smem_gen + STAGES * STAGE_BYTES + 8 * (2 * STAGES + 1));

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int n = blockIdx.z;
    const int n0 = blockIdx.y * BN;
    const int tile_y = blockIdx.x / p.tiles_x, tile_x = blockIdx.x - tile_y * p.tiles_x;
    const int y0 = tile_y * p.th, x0 = tile_x * p.tw;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmA_hi); tma_prefetch_desc(&tmA_lo);
        tma_prefetch_desc(&tmW_hi); tma_prefetch_desc(&tmW_lo);
        for (int s = 0; s < STAGES; s++) { mbar_init(full_bar(s), 1); mbar_init(empty_bar(s), 1); }
        mbar_init(tmem_full_bar, 1);
        mbar_fence_init();
        fence_proxy_async();
    }
    if (warp == 1) tmem_alloc<BN>(tmem_slot);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    pdl_wait();          // PDL: barrier init / TMEM allocation / descriptor prefetch above overlap the predecessor's tail
    const uint32_t tmem_base = *tmem_slot_gen;

    const int num_k = p.k * p.k * p.kc;
    if (warp == 0 && lane == 0) {
        // ---------------- TMA producer
        for (int it = 0; it < num_k; it++) {
            const int s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(empty_bar(s), ph ^ 1);
            mbar_expect_tx(full_bar(s), STAGE_BYTES);
            const int tap = it / p.kc, c0 = (it - tap * p.kc) * 64;
            const int ky = tap / p.k, kx = tap - ky * p.k;
            const uint32_t sa = smem0 + s * STAGE_BYTES;
            const int cx = p.org + x0 * p.stride + kx, cy = p.org + y0 * p.stride + ky;
            tma_load_4d(sa, &tmA_hi, full_bar(s), c0, cx, cy, n);
            tma_load_4d(sa + A_BYTES, &tmA_lo, full_bar(s), c0, cx, cy, n);
            tma_load_3d(sa + 2 * A_BYTES, &tmW_hi, full_bar(s), c0, n0, p.tap_base + tap);
            tma_load_3d(sa + 2 * A_BYTES + W_BYTES, &tmW_lo, full_bar(s), c0, n0, p.tap_base + tap);
        }
    } else if (warp == 1) {
        // ---------------- MMA issuer: whole warp, warp-uniform control flow, one elected lane issues (see tc_conv_halo.cu)
        constexpr uint32_t idesc = make_idesc_bf16(BN, 0, 0);
        const uint64_t dsc = make_desc(0, 16, 1024);
        const bool leader = elect_one();
        for (int it = 0; it < num_k; it++) {
            const int s = it % STAGES, ph = (it / STAGES) & 1;
            mbar_wait(full_bar(s), ph);
            tc_fence_after();
            const uint32_t sa = smem0 + s * STAGE_BYTES;
            if (leader) {
                const uint64_t a_hi = dsc + (sa >> 4), a_lo = dsc + ((sa + A_BYTES) >> 4);
                const uint64_t w_hi = dsc + ((sa + 2 * A_BYTES) >> 4), w_lo = dsc + ((sa + 2 * A_BYTES + W_BYTES) >> 4);
#pragma unroll
                for (int kk = 0; kk < 4; kk++) {
                    mma_bf16(tmem_base, a_lo + 2 * kk, w_hi + 2 * kk, idesc, (it > 0 || kk > 0) ? 1u : 0u);
                    mma_bf16(tmem_base, a_hi + 2 * kk, w_lo + 2 * kk, idesc, 1u);
                    mma_bf16(tmem_base, a_hi + 2 * kk, w_hi + 2 * kk, idesc, 1u);
                }
                mma_commit(empty_bar(s));  // frees the smem stage when these MMAs retire
            }
            __syncwarp();
        }
        if (leader) mma_commit(tmem_full_bar);
    }
    __syncwarp();

    // ---------------- epilogue (all 4 warps)
    mbar_wait(tmem_full_bar, 0);
    tc_fence_after();
    __syncwarp();
    float* red = reinterpret_cast<float*>(smem_gen);  // [4 warps][BN][2], stage memory is free now
    {
        const int r = warp * 32 + lane;
        const int ty = r / p.tw, tx = r - ty * p.tw;
        const int oy = y0 + ty, ox = x0 + tx;
        const int py = oy * p.osy + p.ooy, px = ox * p.osx + p.oox;
        const bool valid = oy < p.ho && ox < p.wo && py < p.OH && px < p.OW;
        float* yrow = p.y + (((long long)n * p.OH + py) * p.OW + px) * p.co + n0;
#pragma unroll 1
        for (int c = 0; c < BN; c += 32) {
            float v[32];
            tmem_ld32(tmem_base + ((uint32_t)(warp * 32) << 16) + (uint32_t)c, v);
            if (p.bias) {
#pragma unroll
                for (int j = 0; j < 32; j++) v[j] += __ldg(p.bias + n0 + c + j);
            }
            if (valid) {
#pragma unroll
                for (int j = 0; j < 32; j += 4)
                    *reinterpret_cast<float4*>(yrow + c + j) = make_float4(v[j], v[j + 1], v[j + 2], v[j + 3]);
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
                red[(warp * BN + c + lane) * 2 + 0] = s1;
                red[(warp * BN + c + lane) * 2 + 1] = s2;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (p.stats) {
        for (int col = threadIdx.x; col < BN; col += 128) {
            float s1 = 0.f, s2 = 0.f;
#pragma unroll
            for (int w = 0; w < 4; w++) { s1 += red[(w * BN + col) * 2]; s2 += red[(w * BN + col) * 2 + 1]; }
            double* dst = p.stats + ((long long)(p.stats_per_n ? n : 0) * p.co + n0 + col) * 2;
            atomicAdd(dst, (double)s1);
            atomicAdd(dst + 1, (double)s2);
        }
    }
    if (warp == 1) {
        tc_fence_after();
        tmem_dealloc<BN>(tmem_base);
    }
}

template <int BN, int STAGES>
static int launch_conv_tc(const CUtensorMap& a_hi, const CUtensorMap& a_lo, const CUtensorMap& w_hi,
                          const CUtensorMap& w_lo, const TcConvP& p, dim3 grid, cudaStream_t st) {
    constexpr int SMEM = STAGES * (2 * A_BYTES + 2 * BN * 128) + 1024 + 256;
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_tc_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, SMEM);
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(conv_tc_kernel<%d>) failed: %s", BN, cudaGetErrorString(e));
            return SKIT_ERR_CUDA;
        }
        attr_set = true;
    }
    launch_pdl(conv_tc_kernel<BN, STAGES>, grid, 128, SMEM, st, a_hi, a_lo, w_hi, w_lo, p);
    return check_launch("conv_tc_kernel");
}

}  // namespace tc

static bool halo_enabled();
int conv_tc_halo_dgrad_full(const skit_operand* d, const void* w_hi, const void* w_lo, int co, int k, float* dx, cudaStream_t st);  // tc_conv_halo.cu

bool conv_tc_eligible(const skit_operand* x, const skit_weights* w, int stride) {
    if (x->fmt != SKIT_FMT_BF16X2 || !w->hi || !w->lo || w->ci != x->c) return false;
    if (stride == 1 && halo_enabled())   // halo kernel: any output width; input channels a multiple of 8, whole 64-chunks or a single thin chunk
        return x->c % 8 == 0 && (x->c % 64 == 0 || x->c < 64);
    return (stride == 1 || stride == 2) && x->c % 64 == 0 && w->co % 64 == 0;
}

int conv_tc_halo_launch(const skit_operand* x, const void* w_hi, const void* w_lo, int ci_pack, int co, int kh, int kw,
                        int ntaps_total, int tap_base, int org, int ho, int wo, const float* bias, float* y,
                        const TcOut* out, double* stats, int stats_mode, cudaStream_t st);  // tc_conv_halo.cu

static bool halo_enabled() {
    static int v = -1;
    if (v < 0) {
        const char* e = getenv("SKIT_TC_HALO");
        v = (e && e[0] == '0') ? 0 : 1;
    }
    return v != 0;
}

// Products per bf16 hi/lo pair in BACKWARD tensor-core launches (input gradients and weight gradients).  The forward convs
// always run the three-term product (A_lo*W_hi + A_hi*W_lo + A_hi*W_hi: ~1e-5 of fp32, the 1e-3 output gate).  Gradients are
// gated at 3e-2 per tensor and the reference's own GPU runs computed them in TF32 (10-bit operands), so by default the
// backward launches drop one cross term: dgrad = (dy_hi + dy_lo) * W_hi, wgrad = (x_hi + x_lo) * dy_hi — a third fewer MMAs,
// half the weight-stage fill — leaving every gradient within ~5e-3 of fp32 (tests/test_baseline_configs_gpu.py prints the
// worst value).  skit_set_backward_terms(3) or SKIT_BWD_TERMS=3 restores the full product.
static int g_bwd_terms = 0;
int bwd_terms() {
    if (g_bwd_terms == 0) {
        const char* e = getenv("SKIT_BWD_TERMS");
        g_bwd_terms = (e && e[0] == '3') ? 3 : 2;
    }
    return g_bwd_terms;
}

// k: taps per side of THIS launch; ntaps_total: taps in the packed filter (weight map extent); tap_base: first tap.
int conv_tc_launch(const skit_operand* x, const void* w_hi, const void* w_lo, int ci, int co, int k, int ntaps_total,
                   int tap_base, int stride, int org, int ho, int wo, const float* bias, float* y, const TcOut* out,
                   double* stats, int stats_mode, cudaStream_t st, int kw = 0) {
    using namespace tc;
    if (kw <= 0) kw = k;
    if (stride == 1 && halo_enabled())   // halo-tile kernel: every tap re-uses one staged activation tile
        return conv_tc_halo_launch(x, w_hi, w_lo, ci, co, k, kw, ntaps_total, tap_base, org, ho, wo, bias, y, out, stats, stats_mode, st);
    if (kw != k) {
        set_error("conv2d_fwd: rectangular (x-folded) filters need the halo-tile kernel (stride 1, SKIT_TC_HALO != 0)");
        return SKIT_ERR_UNSUPPORTED;
    }
    if (!w_lo) {
        set_error("conv_tc: the per-tap kernel runs the three-term product only (two-term launches need the halo-tile kernel)");
        return SKIT_ERR_UNSUPPORTED;
    }
    TcConvP p{};
    p.k = k; p.kc = ci / 64; p.org = org; p.stride = stride; p.tap_base = tap_base; p.ho = ho; p.wo = wo; p.co = co;
    p.tw = (wo <= 8) ? 8 : 16; p.th = 128 / p.tw;
    p.tiles_x = cdiv(wo, p.tw);
    if (out) { p.OH = out->OH; p.OW = out->OW; p.osy = out->osy; p.osx = out->osx; p.ooy = out->ooy; p.oox = out->oox; }
    else { p.OH = ho; p.OW = wo; p.osy = 1; p.osx = 1; p.ooy = 0; p.oox = 0; }
    p.bias = bias; p.y = y; p.stats = stats; p.stats_per_n = stats_mode == SKIT_NORM_INSTANCE;
    const int tiles_y = cdiv(ho, p.th);
    // N tile: the widest tile is the most L2-efficient (the activation box is re-read per N tile), but small maps
    // leave most SMs idle and 148 < CTAs <= 296 wastes a wave.  Cost model per CTA and K step (cycles): the MMAs
    // (3 x 4 x BN/2 at 128x(BN)x16 per ~BN/2 cycles) against the smem fill (A box, 4x for stride 2 because TMA
    // element strides traverse the dense box, plus the weight box) at ~64 B/cycle/SM; total = waves x per-CTA cost.
    int BN = 64;
    {
        const long long tiles = (long long)p.tiles_x * tiles_y * x->n;
        double best = 1e30;
        for (int cand = 256; cand >= 64; cand >>= 1) {
            if (co % cand) continue;
            const double mma = 12.0 * cand * 0.53;
            const double fill = (32768.0 * stride * stride + cand * 256.0) / 64.0 + 300.0;
            const double per = mma > fill ? mma : fill;
            const long long ctas = tiles * (co / cand);
            const double cost = (double)((ctas + 147) / 148) * per;
            if (cost < best * 0.97) { best = cost; BN = cand; }
        }
    }

    CUtensorMap a_hi, a_lo, m_hi, m_lo;
    {
        uint64_t dims[4] = {(uint64_t)ci, (uint64_t)x->wp, (uint64_t)x->hp, (uint64_t)x->n};
        uint64_t strides[3] = {(uint64_t)ci * 2, (uint64_t)ci * 2 * x->wp, (uint64_t)ci * 2 * x->wp * x->hp};
        uint32_t box[4] = {64, (uint32_t)(p.tw * stride), (uint32_t)(p.th * stride), 1};
        uint32_t es[4] = {1, (uint32_t)stride, (uint32_t)stride, 1};
        int rc = encode_bf16_map(&a_hi, x->p0, 4, dims, strides, box, es);
        if (rc) return rc;
        rc = encode_bf16_map(&a_lo, x->p1, 4, dims, strides, box, es);
        if (rc) return rc;
    }
    {
        uint64_t dims[3] = {(uint64_t)ci, (uint64_t)co, (uint64_t)ntaps_total};
        uint64_t strides[2] = {(uint64_t)ci * 2, (uint64_t)ci * 2 * co};
        uint32_t box[3] = {64, (uint32_t)BN, 1};
        int rc = encode_bf16_map(&m_hi, w_hi, 3, dims, strides, box, nullptr);
        if (rc) return rc;
        rc = encode_bf16_map(&m_lo, w_lo, 3, dims, strides, box, nullptr);
        if (rc) return rc;
    }
    dim3 grid(p.tiles_x * tiles_y, co / BN, x->n);
    if (BN == 256) return launch_conv_tc<256, 2>(a_hi, a_lo, m_hi, m_lo, p, grid, st);
    if (BN == 128) return launch_conv_tc<128, 3>(a_hi, a_lo, m_hi, m_lo, p, grid, st);
    return launch_conv_tc<64, 4>(a_hi, a_lo, m_hi, m_lo, p, grid, st);
}

int conv_fwd_tc(const skit_operand* x, const skit_weights* w, int stride, int org, int ho, int wo,
                const float* bias, float* y, double* stats, int stats_mode, cudaStream_t st) {
    const int kw = w->kw > 0 ? w->kw : w->k;
    return conv_tc_launch(x, w->hi, w->lo, x->c, w->co, w->k, w->k * kw, 0, stride, org, ho, wo, bias, y, nullptr,
                          stats, stats_mode, st, kw);
}

int conv_fwd_simt(const skit_operand* x, const skit_weights* w, int stride, int org, int ho, int wo,
                  const float* bias, float* y, double* stats, int stats_mode, cudaStream_t st);
bool conv_head7_eligible(const skit_operand* x, const skit_weights* w, int stride, const double* stats);   // simt_conv.cu
int conv_head7_launch(const skit_operand* x, const skit_weights* w, int org, int ho, int wo, const float* bias, float* y, cudaStream_t st);

}  // namespace skit

using namespace skit;

extern "C" int skit_set_backward_terms(int terms) {
    SKIT_REQUIRE(terms == 2 || terms == 3, "set_backward_terms: 2 or 3");
    g_bwd_terms = terms;
    return SKIT_OK;
}

extern "C" int skit_conv2d_fwd(const skit_operand* x, const skit_weights* w, int stride, int org,
                               int ho, int wo, const float* bias, float* y,
                               double* stats, int stats_mode, int impl, void* stream) {
    SKIT_REQUIRE(x && w && y && x->p0, "conv2d_fwd: null pointer");
    SKIT_REQUIRE(x->fmt == SKIT_FMT_F32 || (x->fmt == SKIT_FMT_BF16X2 && x->p1), "conv2d_fwd: bad operand format");
    SKIT_REQUIRE(w->ci == x->c, "conv2d_fwd: weight ci=%d != operand channels %d", w->ci, x->c);
    SKIT_REQUIRE(stride >= 1 && ho > 0 && wo > 0 && org >= 0, "conv2d_fwd: bad geometry");
    SKIT_REQUIRE(org + (ho - 1) * stride + w->k <= x->hp && org + (wo - 1) * stride + (w->kw > 0 ? w->kw : w->k) <= x->wp,
                 "conv2d_fwd: window exceeds the haloed operand (hp=%d wp=%d k=%d stride=%d ho=%d wo=%d org=%d)",
                 x->hp, x->wp, w->k, stride, ho, wo, org);
    SKIT_REQUIRE(stats == nullptr || stats_mode == SKIT_NORM_INSTANCE || stats_mode == SKIT_NORM_BATCH,
                 "conv2d_fwd: stats given without a stats mode");
    const bool tc_ok = conv_tc_eligible(x, w, stride);
    if (impl == SKIT_IMPL_TC && !tc_ok) {
        set_error("conv2d_fwd: shape not eligible for the tcgen05 path (fmt=%d ci=%d co=%d stride=%d)", x->fmt, x->c, w->co, stride);
        return SKIT_ERR_UNSUPPORTED;
    }
    if (impl == SKIT_IMPL_AUTO && conv_head7_eligible(x, w, stride, stats))    // 7x7 with <= 8 output channels: fp32 pipes beat the MMA waste
        return conv_head7_launch(x, w, org, ho, wo, bias, y, as_stream(stream));
    if (tc_ok && impl != SKIT_IMPL_SIMT) return conv_fwd_tc(x, w, stride, org, ho, wo, bias, y, stats, stats_mode, as_stream(stream));
    SKIT_REQUIRE(w->f32, "conv2d_fwd: CUDA-core path needs the fp32 weight pack");
    return conv_fwd_simt(x, w, stride, org, ho, wo, bias, y, stats, stats_mode, as_stream(stream));
}

// Stride-2 input gradient on the tensor cores: the four output parities (iy%2, ix%2) are four independent
// stride-1 convolutions of the zero-haloed output gradient with (k/2 x k/2) sub-filters (mode-3 pack),
// each writing its own interleaved quarter of dx.
extern "C" int skit_conv2d_dgrad_s2(const skit_operand* dy, int dy_pad, const skit_weights* wp, int k,
                                    int ho, int wo, int hp, int wp_, float* dx, void* stream) {
    SKIT_REQUIRE(dy && wp && dx && dy->p0 && dy->p1 && wp->hi && wp->lo, "conv2d_dgrad_s2: null pointer");
    SKIT_REQUIRE(dy->fmt == SKIT_FMT_BF16X2 && k % 2 == 0 && dy_pad == k / 2 - 1, "conv2d_dgrad_s2: needs a bf16x2 dy operand with halo k/2-1");
    SKIT_REQUIRE(dy->c % 64 == 0 && wp->ci == dy->c && (wp->co % 64 == 0 || halo_enabled()),
                 "conv2d_dgrad_s2: the gradient's channels must be a multiple of 64 (and so the input's, without the halo kernel)");
    SKIT_REQUIRE(dy->hp == ho + 2 * dy_pad && dy->wp == wo + 2 * dy_pad, "conv2d_dgrad_s2: dy operand dims mismatch");
    const int kh = k / 2;
    for (int py = 0; py < 2; py++)
        for (int px = 0; px < 2; px++) {
            const int A = (hp - py + 1) / 2, B = (wp_ - px + 1) / 2;  // outputs of this parity
            if (A <= 0 || B <= 0) continue;
            TcOut out{hp, wp_, 2, 2, py, px};
            int rc = conv_tc_launch(dy, wp->hi, (bwd_terms() == 2 && halo_enabled()) ? nullptr : wp->lo, dy->c, wp->co, kh, 4 * kh * kh, (py * 2 + px) * kh * kh, 1, 0, A, B,
                                    nullptr, dx, &out, nullptr, SKIT_NORM_NONE, as_stream(stream));
            if (rc) return rc;
        }
    return SKIT_OK;
}


// Stride-1 input gradient on the tensor cores: dx = full correlation of the zero-haloed (k-1) gradient operand with the
// mode-1 (flipped, transposed) pack.  With the halo-tile kernel the border strips run as cheap extra regions of the same grid.
extern "C" int skit_conv2d_dgrad_s1(const skit_operand* dy, const skit_weights* w1, float* dx, void* stream) {
    SKIT_REQUIRE(dy && w1 && dx && dy->p0, "conv2d_dgrad_s1: null pointer");
    const int k = w1->k;
    SKIT_REQUIRE(w1->ci == dy->c && dy->hp >= k && dy->wp >= k, "conv2d_dgrad_s1: pack / operand mismatch");
    const int H = dy->hp - k + 1, W = dy->wp - k + 1;
    if (halo_enabled() && dy->fmt == SKIT_FMT_BF16X2 && w1->hi && w1->lo && w1->kw == 0 && dy->c % 64 == 0)
        return conv_tc_halo_dgrad_full(dy, w1->hi, bwd_terms() == 2 ? nullptr : w1->lo, w1->co, k, dx, as_stream(stream));
    return skit_conv2d_fwd(dy, w1, 1, 0, H, W, nullptr, dx, nullptr, SKIT_NORM_NONE, SKIT_IMPL_AUTO, stream);
}
