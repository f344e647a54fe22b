// Generic fp32 CUDA-core implicit-GEMM convolution kernels (forward, gather dgrad, wgrad) and the
// weight (un)packers.  These cover every layer shape of the path (tiny channel counts, strides,
// first/last 7x7 convs); the tcgen05 kernel in tc_conv.cu takes the tensor-core-eligible layers.
#include <cstdlib>
#include "skit_common.cuh"

namespace skit {

// Packed fp32 FMA (Blackwell FFMA2): two independent FMAs per lane per instruction on 64-bit register pairs.  The three-register
// FFMA issues every second cycle per scheduler; in the FMA-bound 7x7 head kernels the channel-quad dot product runs as two partial
// sums (x,y lanes), added once at the end.
__device__ __forceinline__ float2 ffma2(float2 a, float2 b, float2 c) {
    unsigned long long rd;
    asm("fma.rn.f32x2 %0, %1, %2, %3;"
        : "=l"(rd)
        : "l"(*reinterpret_cast<unsigned long long*>(&a)), "l"(*reinterpret_cast<unsigned long long*>(&b)),
          "l"(*reinterpret_cast<unsigned long long*>(&c)));
    return *reinterpret_cast<float2*>(&rd);
}


struct ConvP {
    const float* x0;
    const __nv_bfloat16* xh;
    const __nv_bfloat16* xl;
    int hp, wp, ci;       // operand (A side) geometry
    const float* w;       // [K][ncol]
    const float* bias;
    float* y;
    double* stats;
    int stats_per_n;      // 1: stats index = n (instance), 0: index 0 (batch)
    int k, stride, org;
    int ho, wo;           // fwd: output size; dgrad: dy size
    int ncol;             // fwd: co; dgrad: ci
    int K;                // reduction length
    int M;                // rows per image (fwd: ho*wo; dgrad: hp*wp)
    int arows;            // dgrad: channels of dy (co)
    int flat;             // 1: rows run over the whole batch (n*M), grid.z == 1 (batch statistics only)
    int n_img;
    int crop;             // dgrad as a transposed conv: output row m maps to padded coordinate (y + crop, x + crop)
    int phased;           // MODE 1, stride > 1: grid.z = image * stride^2 + parity class; a block's rows share the parity of
                          // (y + crop, x + crop), so only the (k/stride)^2 taps that can hit an output sample are visited
    int yc, yoff;         // output channel stride / offset (0: dense ncol) — writes a channel slice of a wider map
};

constexpr int BM = 128, BN = 64, BK = 16;
// 16-wide column tiles when that covers the columns with at most half the padding of 64-wide ones
static inline bool narrow_tile(int ncol) { return ((ncol + 15) / 16) * 16 * 2 <= ((ncol + 63) / 64) * 64; }

// MODE 0: forward valid conv.  MODE 1: gather dgrad.  CPT = output columns per thread: the tile is 128 rows x 16*CPT
// columns (64 wide by default; 16 wide for the default U-Net's thin layers, whose 2..20 output channels would leave a
// 64-wide tile mostly empty).
template <int MODE, int FMT, int CPT = 4>
__global__ void __launch_bounds__(256) conv_simt_kernel(ConvP p) {
    constexpr int BN = 16 * CPT;
    __shared__ float As[BK][BM + 4];
    __shared__ float Bs[BK][BN + 4];
    const int tid = threadIdx.x;
    int nz = blockIdx.z;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int kk = tid & 15, rg = tid >> 4;
    const int ty = tid >> 4, tx = tid & 15;
    long long Mtot = p.flat ? (long long)p.n_img * p.M : p.M;
    // parity-class geometry (phased gather dgrad): class (py, px) of (y + crop, x + crop) mod stride
    int ph_py = 0, ph_px = 0, ph_y0 = 0, ph_x0 = 0, ph_nx = 1, ph_kty = 1, ph_ktx = 1, Keff = p.K;
    if (MODE == 1 && p.phased) {
        const int s = p.stride, cls = blockIdx.z % (s * s);
        nz = blockIdx.z / (s * s);
        ph_py = cls / s; ph_px = cls - ph_py * s;
        ph_y0 = ((ph_py - p.crop) % s + s) % s;      // first output row of this class
        ph_x0 = ((ph_px - p.crop) % s + s) % s;
        const int ny = p.hp > ph_y0 ? (p.hp - ph_y0 + s - 1) / s : 0;
        ph_nx = p.wp > ph_x0 ? (p.wp - ph_x0 + s - 1) / s : 0;
        Mtot = (long long)ny * ph_nx;
        ph_kty = (p.k - ph_py + s - 1) / s; ph_ktx = (p.k - ph_px + s - 1) / s;
        Keff = ph_kty * ph_ktx * p.arows;
        if (m0 >= Mtot) return;
    }

    long long base[8];
    int ry[8], rx[8], rn[8];
    bool rvalid[8];
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int m = m0 + rg + 16 * i;
        rvalid[i] = m < Mtot;
        int n = nz;
        if (p.flat) { n = m / p.M; m -= n * p.M; }
        rn[i] = n;
        if (MODE == 0) {
            int oy = m / p.wo, ox = m - oy * p.wo;
            base[i] = (((long long)n * p.hp + p.org + oy * p.stride) * p.wp + p.org + ox * p.stride) * p.ci;
            ry[i] = 0; rx[i] = 0;
        } else if (p.phased) {
            const int jy = ph_nx ? m / ph_nx : 0, jx = m - jy * ph_nx;
            ry[i] = ph_y0 + jy * p.stride + p.crop; rx[i] = ph_x0 + jx * p.stride + p.crop;
            base[i] = 0;
        } else {
            int iy = m / p.wp, ix = m - iy * p.wp;
            ry[i] = iy + p.crop; rx[i] = ix + p.crop;
            base[i] = 0;
        }
    }
    float2 acc2[4][CPT];        // rows (2i, 2i+1) of column j: packed fp32 FMAs, same summation order as scalar FMAs
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < CPT; j++) acc2[i][j] = make_float2(0.f, 0.f);

    const int brow = tid >> 4, bcol = (tid & 15) * CPT;

    for (int k0 = 0; k0 < Keff; k0 += BK) {
        const int kg = k0 + kk;
        const bool kvalid = kg < Keff;
        if (MODE == 0) {
            int tap = kg / p.ci, c = kg - tap * p.ci;
            int ky = tap / p.k, kx = tap - ky * p.k;
            long long off = ((long long)ky * p.wp + kx) * p.ci + c;
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float v = 0.f;
                if (kvalid && rvalid[i]) {
                    if (FMT == SKIT_FMT_F32) v = __ldg(p.x0 + base[i] + off);
                    else v = __bfloat162float(p.xh[base[i] + off]) + __bfloat162float(p.xl[base[i] + off]);
                }
                As[kk][rg + 16 * i] = v;
            }
        } else {
            int tap = kg / p.arows, o = kg - tap * p.arows;
            int ky = tap / p.k, kx = tap - ky * p.k;
            if (p.phased) {     // tap = index among the class's valid taps
                const int tyi = tap / ph_ktx, txi = tap - tyi * ph_ktx;
                ky = ph_py + tyi * p.stride; kx = ph_px + txi * p.stride;
            }
#pragma unroll
            for (int i = 0; i < 8; i++) {
                float v = 0.f;
                if (kvalid && rvalid[i]) {
                    int ty_ = ry[i] - ky, tx_ = rx[i] - kx;
                    if (ty_ >= 0 && tx_ >= 0) {
                        int oy = ty_ / p.stride, ox = tx_ / p.stride;
                        if (oy * p.stride == ty_ && ox * p.stride == tx_ && oy < p.ho && ox < p.wo)
                            v = __ldg(p.x0 + (((long long)rn[i] * p.ho + oy) * p.wo + ox) * p.arows + o);
                    }
                }
                As[kk][rg + 16 * i] = v;
            }
        }
        {
            const int kgb = k0 + brow;
            long long wrow = kgb;
            if (MODE == 1 && p.phased) {    // row of the [(ky*k + kx)*arows + o][ncol] pack for this class-local K index
                const int tap = kgb / p.arows, o = kgb - tap * p.arows;
                const int tyi = tap / ph_ktx, txi = tap - tyi * ph_ktx;
                wrow = (long long)((ph_py + tyi * p.stride) * p.k + ph_px + txi * p.stride) * p.arows + o;
            }
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                int col = n0 + bcol + j;
                Bs[brow][bcol + j] = (kgb < Keff && col < p.ncol) ? __ldg(p.w + wrow * p.ncol + col) : 0.f;
            }
        }
        __syncthreads();
#pragma unroll
        for (int k2 = 0; k2 < BK; k2++) {
            float b[CPT];
            const float4 a0 = *reinterpret_cast<const float4*>(&As[k2][ty * 8]);
            const float4 a1 = *reinterpret_cast<const float4*>(&As[k2][ty * 8 + 4]);
            const float2 ap[4] = {make_float2(a0.x, a0.y), make_float2(a0.z, a0.w), make_float2(a1.x, a1.y), make_float2(a1.z, a1.w)};
#pragma unroll
            for (int j = 0; j < CPT; j++) b[j] = Bs[k2][tx * CPT + j];
#pragma unroll
            for (int j = 0; j < CPT; j++) {
                const float2 bb = make_float2(b[j], b[j]);
#pragma unroll
                for (int i = 0; i < 4; i++) acc2[i][j] = ffma2(ap[i], bb, acc2[i][j]);
            }
        }
        __syncthreads();
    }
    float acc[8][CPT];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < CPT; j++) { acc[2 * i][j] = acc2[i][j].x; acc[2 * i + 1][j] = acc2[i][j].y; }

    float s[CPT], q[CPT];
#pragma unroll
    for (int j = 0; j < CPT; j++) { s[j] = 0.f; q[j] = 0.f; }
#pragma unroll
    for (int i = 0; i < 8; i++) {
        int m = m0 + ty * 8 + i;
        if (m >= Mtot) continue;
        long long mout = m;
        if (MODE == 1 && p.phased) {    // class-local row -> position in the full output map
            const int jy = m / ph_nx, jx = m - jy * ph_nx;
            mout = (long long)(ph_y0 + jy * p.stride) * p.wp + ph_x0 + jx * p.stride;
        }
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            int col = n0 + tx * CPT + j;
            if (col >= p.ncol) continue;
            float v = acc[i][j] + (p.bias ? p.bias[col] : 0.f);
            p.y[((long long)(p.flat ? 0 : nz) * p.M + mout) * (p.yc ? p.yc : p.ncol) + p.yoff + col] = v;
            s[j] += v; q[j] += v * v;
        }
    }
    if (p.stats) {
        float* red = &As[0][0];  // 2*BN floats
        if (tid < 2 * BN) red[tid] = 0.f;
        __syncthreads();
#pragma unroll
        for (int j = 0; j < CPT; j++) {
            atomicAdd(&red[tx * CPT + j], s[j]);
            atomicAdd(&red[BN + tx * CPT + j], q[j]);
        }
        __syncthreads();
        if (tid < BN && n0 + tid < p.ncol) {
            double* dst = p.stats + ((long long)(p.stats_per_n ? nz : 0) * p.ncol + n0 + tid) * 2;
            atomicAdd(dst, (double)red[tid]);
            atomicAdd(dst + 1, (double)red[BN + tid]);
        }
    }
}

// ---------------------------------------------------------------------------------- thin-N convolution
// ncol <= 8 (generator head 64->5, PatchGAN head 512->1, input gradients of the 4/7-channel PatchGAN
// stems): a 128x64 GEMM tile would waste >90% of its columns, so one warp owns PIX output rows,
// its lanes stride over the reduction (coalesced 128 B channel runs), and the NOUT x PIX partial
// sums are combined with warp shuffles.  Requires the per-tap channel count to be a multiple of 32.
template <int MODE, int FMT, int NOUT, int PIX>
__global__ void __launch_bounds__(256) conv_thin_kernel(ConvP p, int n_img, int groups_per_row) {
    const int lane = threadIdx.x & 31;
    const long long warp = ((long long)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
    const int rows_y = (MODE == 0) ? p.ho : p.hp;
    const long long total = (long long)n_img * rows_y * groups_per_row;
    if (warp >= total) return;
    const int gx = (int)(warp % groups_per_row);
    long long t = warp / groups_per_row;
    const int y = (int)(t % rows_y);
    const int n = (int)(t / rows_y);
    const int x0 = gx * PIX;
    const int row_w = (MODE == 0) ? p.wo : p.wp;
    const int ch = (MODE == 0) ? p.ci : p.arows;  // channels per tap on the reduction side

    float acc[PIX][NOUT];
#pragma unroll
    for (int i = 0; i < PIX; i++)
#pragma unroll
        for (int j = 0; j < NOUT; j++) acc[i][j] = 0.f;

    // input-gradient mode visits only the taps that can hit an output sample: (y - ky) % stride == 0 (and the same in x
    // when the warp owns a single column) — a stride-2 k4 layer has 4 such taps per pixel, not 16
    const int ky0 = (MODE == 1) ? (y % p.stride) : 0, kys = (MODE == 1) ? p.stride : 1;
    const int kx0 = (MODE == 1 && PIX == 1) ? (x0 % p.stride) : 0, kxs = (MODE == 1 && PIX == 1) ? p.stride : 1;
    for (int ky = ky0; ky < p.k; ky += kys) {
        for (int kx = kx0; kx < p.k; kx += kxs) {
            long long abase[PIX];
            bool av[PIX];
#pragma unroll
            for (int i = 0; i < PIX; i++) {
                const int x = x0 + i;
                av[i] = x < row_w;
                if (MODE == 0) {
                    abase[i] = (((long long)n * p.hp + p.org + y * p.stride + ky) * p.wp + p.org + x * p.stride + kx) * p.ci;
                } else {
                    const int ty_ = y - ky, tx_ = x - kx;
                    const int oy = ty_ / p.stride, ox = tx_ / p.stride;
                    av[i] = av[i] && ty_ >= 0 && tx_ >= 0 && oy * p.stride == ty_ && ox * p.stride == tx_ && oy < p.ho && ox < p.wo;
                    abase[i] = (((long long)n * p.ho + oy) * p.wo + ox) * p.arows;
                }
            }
            const float* wtap = p.w + (long long)(ky * p.k + kx) * ch * p.ncol;
            for (int c0 = 0; c0 < ch; c0 += 32) {
                const int c = c0 + lane;
                float wv[NOUT];
#pragma unroll
                for (int j = 0; j < NOUT; j++) wv[j] = (j < p.ncol) ? __ldg(wtap + (long long)c * p.ncol + j) : 0.f;
#pragma unroll
                for (int i = 0; i < PIX; i++) {
                    float a = 0.f;
                    if (av[i]) {
                        if (FMT == SKIT_FMT_F32) a = __ldg(p.x0 + abase[i] + c);
                        else a = __bfloat162float(p.xh[abase[i] + c]) + __bfloat162float(p.xl[abase[i] + c]);
                    }
#pragma unroll
                    for (int j = 0; j < NOUT; j++) acc[i][j] = fmaf(a, wv[j], acc[i][j]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < PIX; i++)
#pragma unroll
        for (int j = 0; j < NOUT; j++) acc[i][j] = warp_sum(acc[i][j]);
    if (lane == 0) {
#pragma unroll
        for (int i = 0; i < PIX; i++) {
            const int x = x0 + i;
            if (x >= row_w) continue;
            float* dst = p.y + (((long long)n * rows_y + y) * row_w + x) * p.ncol;
#pragma unroll
            for (int j = 0; j < NOUT; j++)
                if (j < p.ncol) dst[j] = acc[i][j] + (p.bias ? p.bias[j] : 0.f);
        }
    }
}

template <int MODE, int FMT, int PIX>
static int launch_thin_pix(const ConvP& p, int n_img, cudaStream_t st) {
    const int row_w = (MODE == 0) ? p.wo : p.wp, rows_y = (MODE == 0) ? p.ho : p.hp;
    const int gpr = cdiv(row_w, PIX);
    const long long warps = (long long)n_img * rows_y * gpr;
    const int blocks = (int)cdivll(warps, 8);
    if (p.ncol <= 1) conv_thin_kernel<MODE, FMT, 1, PIX><<<blocks, 256, 0, st>>>(p, n_img, gpr);
    else if (p.ncol <= 4) conv_thin_kernel<MODE, FMT, 4, PIX><<<blocks, 256, 0, st>>>(p, n_img, gpr);
    else conv_thin_kernel<MODE, FMT, 8, PIX><<<blocks, 256, 0, st>>>(p, n_img, gpr);
    return check_launch("conv_thin_kernel");
}

// Pixels per warp: 4 for forward convs and for stride-1 input gradients (weights and the shuffle reduction amortised over
// four outputs); strided input gradients keep one pixel per warp so that both tap parities can be skipped.
template <int MODE, int FMT>
static int launch_thin(const ConvP& p, int n_img, cudaStream_t st) {
    if (MODE == 0 || p.stride == 1) return launch_thin_pix<MODE, FMT, 4>(p, n_img, st);
    return launch_thin_pix<MODE, FMT, 1>(p, n_img, st);
}

// ---------------------------------------------------------------------------------- thin-output 7x7 head
// Generator head: Conv2d(64 -> 5, k7) over the reflect-padded activation (networks.py:1124-1126).  On the tensor cores this
// layer wastes >95 % of every MMA (5 useful columns) and is bound by the per-MMA operand fetch (0.55 ms at 512x512); its
// 8.2 GFLOP fit the fp32 pipes better.  A block owns a 32x32 output tile; a thread owns the 2x2 pixels (tx + 16a, ty + 16b)
// so that a warp's shared-memory reads are consecutive 16-byte pixel quads (conflict free), and walks the 64 input channels in
// chunks of 8: the (32+6)^2 x 8 input tile and the chunk's 49 x 8 x CO weights sit in shared memory as channel quads.  Per
// (tap, quad): 4 input float4 + CO broadcast weight float4 feed 4*CO*4 FMAs — FMA bound, not LDS bound.
// Operand: fp32 or bf16 hi/lo; weights: the tcgen05 pack (bf16 hi/lo [tap][co][ci]) or nullptr planes -> fp32 pack.
template <int CO>
__global__ void __launch_bounds__(256) conv_head7_kernel(const float* __restrict__ x0, const __nv_bfloat16* __restrict__ xh,
                                                         const __nv_bfloat16* __restrict__ xl, int fmt, int hp, int wp, int ci, int org,
                                                         const __nv_bfloat16* __restrict__ wh, const __nv_bfloat16* __restrict__ wl, int wci,
                                                         const float* __restrict__ bias, float* __restrict__ y, int ho, int wo, int co) {
    constexpr int T = 32, HT = T + 6, CH = 8;
    extern __shared__ float4 sm4[];
    float4* sx = sm4;                              // [2 quads][HT][HT]
    float4* sw = sm4 + 2 * HT * HT;                // [49 taps][2 quads][CO]
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    const int n = blockIdx.z;
    const int x_0 = blockIdx.x * T, y_0 = blockIdx.y * T;
    float2 acc[2][2][CO];       // two partial sums per output (FFMA2)
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++)
#pragma unroll
            for (int o = 0; o < CO; o++) acc[a][b][o] = make_float2(0.f, 0.f);

    for (int c0 = 0; c0 < ci; c0 += CH) {
        // stage the input tile: one (pixel, quad) per thread-iteration
        for (int i = threadIdx.x; i < 2 * HT * HT; i += 256) {
            const int q = i / (HT * HT), r = i - q * HT * HT;
            const int py = r / HT, px = r - py * HT;
            const int gy = org + y_0 + py, gx = org + x_0 + px;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gy < hp && gx < wp) {
                const long long a = (((long long)n * hp + gy) * wp + gx) * ci + c0 + q * 4;
                if (fmt == SKIT_FMT_F32) v = *reinterpret_cast<const float4*>(x0 + a);
                else {
                    const uint2 h = *reinterpret_cast<const uint2*>(xh + a), l = *reinterpret_cast<const uint2*>(xl + a);
                    const __nv_bfloat162 h0 = *reinterpret_cast<const __nv_bfloat162*>(&h.x), h1 = *reinterpret_cast<const __nv_bfloat162*>(&h.y);
                    const __nv_bfloat162 l0 = *reinterpret_cast<const __nv_bfloat162*>(&l.x), l1 = *reinterpret_cast<const __nv_bfloat162*>(&l.y);
                    v.x = __bfloat162float(h0.x) + __bfloat162float(l0.x); v.y = __bfloat162float(h0.y) + __bfloat162float(l0.y);
                    v.z = __bfloat162float(h1.x) + __bfloat162float(l1.x); v.w = __bfloat162float(h1.y) + __bfloat162float(l1.y);
                }
            }
            sx[i] = v;
        }
        // stage the chunk's weights: sw[(tap*2 + q)*CO + o] = w[tap][o][c0 + 4q .. +3]
        for (int i = threadIdx.x; i < 49 * 2 * CO; i += 256) {
            const int o = i % CO, t = i / CO;
            const int q = t & 1, tap = t >> 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (o < co) {
                const long long a = ((long long)tap * co + o) * wci + c0 + q * 4;
                float f[4];
#pragma unroll
                for (int j = 0; j < 4; j++) f[j] = __bfloat162float(wh[a + j]) + __bfloat162float(wl[a + j]);
                v = make_float4(f[0], f[1], f[2], f[3]);
            }
            sw[i] = v;
        }
        __syncthreads();
#pragma unroll 1
        for (int ky = 0; ky < 7; ky++) {
#pragma unroll
            for (int kx = 0; kx < 7; kx++) {
#pragma unroll
                for (int q = 0; q < 2; q++) {
                    float4 in[2][2];
#pragma unroll
                    for (int a = 0; a < 2; a++)
#pragma unroll
                        for (int b = 0; b < 2; b++)
                            in[a][b] = sx[(q * HT + ty + 16 * b + ky) * HT + tx + 16 * a + kx];
                    const float4* wq = sw + ((ky * 7 + kx) * 2 + q) * CO;
#pragma unroll
                    for (int o = 0; o < CO; o++) {
                        const float4 w4 = wq[o];
#pragma unroll
                        for (int a = 0; a < 2; a++)
#pragma unroll
                            for (int b = 0; b < 2; b++) {
                                float2 t = acc[a][b][o];
                                t = ffma2(make_float2(in[a][b].x, in[a][b].y), make_float2(w4.x, w4.y), t);
                                t = ffma2(make_float2(in[a][b].z, in[a][b].w), make_float2(w4.z, w4.w), t);
                                acc[a][b][o] = t;
                            }
                    }
                }
            }
        }
        __syncthreads();
    }
#pragma unroll
    for (int a = 0; a < 2; a++)
#pragma unroll
        for (int b = 0; b < 2; b++) {
            const int ox = x_0 + tx + 16 * a, oy = y_0 + ty + 16 * b;
            if (ox >= wo || oy >= ho) continue;
            float* dst = y + (((long long)n * ho + oy) * wo + ox) * co;
#pragma unroll
            for (int o = 0; o < CO; o++)
                if (o < co) dst[o] = (acc[a][b][o].x + acc[a][b][o].y) + (bias ? bias[o] : 0.f);
        }
}


// Second formulation, used for a SINGLE output channel (CO = 1: 72 registers): a thread owns 8 CONSECUTIVE pixels of one output row, so the 7 taps of a filter row share a
// sliding window of 14 input quads (14 + 7*CO shared-memory loads feed 7*8*CO*4 FMAs per (filter row, channel quad): ratio 1:23 for
// CO = 5 against 1:9 above — a 16-byte shared-memory load costs a warp 4 cycles of the LDS pipe, an FMA a quarter cycle of issue,
// so anything below 1:16 is LDS bound).  Block = 64 x 32 output pixels (8 x 32 threads); the input tile (70 x 38 pixels, 8
// channels) is stored with one quad of skew per 8 pixels so that the 8 threads of a row, whose windows start 8 pixels apart,
// read 8 different bank groups.
template <int CO>
__global__ void __launch_bounds__(256) conv_head7_row8_kernel(const float* __restrict__ x0, const __nv_bfloat16* __restrict__ xh,
                                                              const __nv_bfloat16* __restrict__ xl, int fmt, int hp, int wp, int ci, int org,
                                                              const __nv_bfloat16* __restrict__ wh, const __nv_bfloat16* __restrict__ wl, int wci,
                                                              const float* __restrict__ bias, float* __restrict__ y, int ho, int wo, int co) {
    constexpr int TX = 64, TYH = 32, HTX = TX + 6, HTY = TYH + 6, PITCH = 80, CH = 8;
    extern __shared__ float4 sm4[];
    float4* sx = sm4;                              // [2 quads][HTY][PITCH], pixel px at px + (px >> 3)
    float4* sw = sm4 + 2 * HTY * PITCH;            // [49 taps][2 quads][CO]
    const int tx = threadIdx.x & 7, ty = threadIdx.x >> 3;
    const int n = blockIdx.z;
    const int x_0 = blockIdx.x * TX, y_0 = blockIdx.y * TYH;
    float2 acc[8][CO];          // two partial sums per output (FFMA2)
#pragma unroll
    for (int j = 0; j < 8; j++)
#pragma unroll
        for (int o = 0; o < CO; o++) acc[j][o] = make_float2(0.f, 0.f);

    for (int c0 = 0; c0 < ci; c0 += CH) {
        for (int i = threadIdx.x; i < 2 * HTY * HTX; i += 256) {
            const int q = i / (HTY * HTX), r = i - q * HTY * HTX;
            const int py = r / HTX, px = r - py * HTX;
            const int gy = org + y_0 + py, gx = org + x_0 + px;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (gy < hp && gx < wp) {
                const long long a = (((long long)n * hp + gy) * wp + gx) * ci + c0 + q * 4;
                if (fmt == SKIT_FMT_F32) v = *reinterpret_cast<const float4*>(x0 + a);
                else {
                    const uint2 h = *reinterpret_cast<const uint2*>(xh + a), l = *reinterpret_cast<const uint2*>(xl + a);
                    const __nv_bfloat162 h0 = *reinterpret_cast<const __nv_bfloat162*>(&h.x), h1 = *reinterpret_cast<const __nv_bfloat162*>(&h.y);
                    const __nv_bfloat162 l0 = *reinterpret_cast<const __nv_bfloat162*>(&l.x), l1 = *reinterpret_cast<const __nv_bfloat162*>(&l.y);
                    v.x = __bfloat162float(h0.x) + __bfloat162float(l0.x); v.y = __bfloat162float(h0.y) + __bfloat162float(l0.y);
                    v.z = __bfloat162float(h1.x) + __bfloat162float(l1.x); v.w = __bfloat162float(h1.y) + __bfloat162float(l1.y);
                }
            }
            sx[(q * HTY + py) * PITCH + px + (px >> 3)] = v;
        }
        for (int i = threadIdx.x; i < 49 * 2 * CO; i += 256) {
            const int o = i % CO, t = i / CO;
            const int q = t & 1, tap = t >> 1;
            float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
            if (o < co) {
                const long long a = ((long long)tap * co + o) * wci + c0 + q * 4;
                float f[4];
#pragma unroll
                for (int j = 0; j < 4; j++) f[j] = __bfloat162float(wh[a + j]) + __bfloat162float(wl[a + j]);
                v = make_float4(f[0], f[1], f[2], f[3]);
            }
            sw[i] = v;
        }
        __syncthreads();
#pragma unroll 1
        for (int ky = 0; ky < 7; ky++) {
#pragma unroll
            for (int q = 0; q < 2; q++) {
                float4 in[14];
                const float4* row = sx + (q * HTY + ty + ky) * PITCH + tx * 9;
#pragma unroll
                for (int i = 0; i < 14; i++) in[i] = row[i + (i >> 3)];
#pragma unroll
                for (int kx = 0; kx < 7; kx++) {
                    const float4* wq = sw + ((ky * 7 + kx) * 2 + q) * CO;
#pragma unroll
                    for (int o = 0; o < CO; o++) {
                        const float4 w4 = wq[o];
#pragma unroll
                        for (int j = 0; j < 8; j++) {
                            float2 t = acc[j][o];
                            t = ffma2(make_float2(in[j + kx].x, in[j + kx].y), make_float2(w4.x, w4.y), t);
                            t = ffma2(make_float2(in[j + kx].z, in[j + kx].w), make_float2(w4.z, w4.w), t);
                            acc[j][o] = t;
                        }
                    }
                }
            }
        }
        __syncthreads();
    }
    const int oy = y_0 + ty;
    if (oy >= ho) return;
#pragma unroll
    for (int j = 0; j < 8; j++) {
        const int ox = x_0 + tx * 8 + j;
        if (ox >= wo) continue;
        float* dst = y + (((long long)n * ho + oy) * wo + ox) * co;
#pragma unroll
        for (int o = 0; o < CO; o++)
            if (o < co) dst[o] = (acc[j][o].x + acc[j][o].y) + (bias ? bias[o] : 0.f);
    }
}

bool conv_head7_eligible(const skit_operand* x, const skit_weights* w, int stride, const double* stats) {
    return w->k == 7 && w->kw == 0 && stride == 1 && w->co <= 8 && !stats && w->hi && w->lo && x->c % 8 == 0 && w->ci == x->c &&
           (x->fmt == SKIT_FMT_F32 || x->fmt == SKIT_FMT_BF16X2);
}

int conv_head7_launch(const skit_operand* x, const skit_weights* w, int org, int ho, int wo, const float* bias, float* y, cudaStream_t st) {
    constexpr int T = 32, HT = T + 6;
    const size_t smem = (size_t)(2 * HT * HT + 49 * 2 * 8) * sizeof(float4);
    static bool attr_set = false;
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(conv_head7_kernel<8>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_head7_kernel<5>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_head7_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e == cudaSuccess) e = cudaFuncSetAttribute(conv_head7_row8_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)((2 * 38 * 80 + 49 * 2) * sizeof(float4)));
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(conv_head7_kernel) failed: %s", cudaGetErrorString(e));
            return SKIT_ERR_CUDA;
        }
        attr_set = true;
    }
    dim3 grid(cdiv(wo, T), cdiv(ho, T), x->n);
    const __nv_bfloat16 *wh = (const __nv_bfloat16*)w->hi, *wl = (const __nv_bfloat16*)w->lo;
    // (an 8-consecutive-pixels-per-thread formulation with a sliding input window was measured SLOWER: 771 vs 578 us at
    //  768x768, 128 registers and half the resident warps; the 2x2 interleaved mapping below stays)
    static int row8 = -1;       // SKIT_HEAD7_ROW8=0: the 2x2 mapping for the single-channel case too (A/B timing)
    if (row8 < 0) { const char* e = getenv("SKIT_HEAD7_ROW8"); row8 = (e && e[0] == '0') ? 0 : 1; }
    if (w->co == 1 && row8) {   // single output channel (the stem's input gradient w.r.t. the sketch channel, PatchNCE query branch)
        dim3 grid8(cdiv(wo, 64), cdiv(ho, 32), x->n);
        conv_head7_row8_kernel<1><<<grid8, 256, (2 * 38 * 80 + 49 * 2) * sizeof(float4), st>>>(
            (const float*)x->p0, (const __nv_bfloat16*)x->p0, (const __nv_bfloat16*)x->p1, x->fmt, x->hp, x->wp, x->c, org, wh, wl, w->ci, bias, y, ho, wo, w->co);
        return check_launch("conv_head7_row8_kernel");
    }
    if (w->co == 1)
        conv_head7_kernel<1><<<grid, 256, smem, st>>>((const float*)x->p0, (const __nv_bfloat16*)x->p0, (const __nv_bfloat16*)x->p1, x->fmt,
                                                     x->hp, x->wp, x->c, org, wh, wl, w->ci, bias, y, ho, wo, w->co);
    else if (w->co <= 5)
        conv_head7_kernel<5><<<grid, 256, smem, st>>>((const float*)x->p0, (const __nv_bfloat16*)x->p0, (const __nv_bfloat16*)x->p1, x->fmt,
                                                     x->hp, x->wp, x->c, org, wh, wl, w->ci, bias, y, ho, wo, w->co);
    else
        conv_head7_kernel<8><<<grid, 256, smem, st>>>((const float*)x->p0, (const __nv_bfloat16*)x->p0, (const __nv_bfloat16*)x->p1, x->fmt,
                                                     x->hp, x->wp, x->c, org, wh, wl, w->ci, bias, y, ho, wo, w->co);
    return check_launch("conv_head7_kernel");
}

// ---------------------------------------------------------------------------------- wgrad
struct WgradP {
    const float* x0; const __nv_bfloat16* xh; const __nv_bfloat16* xl;
    int hp, wp, ci, org;
    const float* d0; const __nv_bfloat16* dh; const __nv_bfloat16* dl;
    int dhp, dwp, co, dorg;
    int k, stride, ho, wo;
    int Kf;        // k*k*ci
    int chunk;     // pixels per split
    int splits;    // per image
    float* dwf;
};

constexpr int WM = 64, WN = 64, WK = 16;

template <int XFMT, int DFMT>
__global__ void __launch_bounds__(256) wgrad_simt_kernel(WgradP p) {
    __shared__ float As[WK][WM + 4];  // [pix][o]
    __shared__ float Bs[WK][WN + 4];  // [pix][col]
    const int tid = threadIdx.x;
    const int n = blockIdx.z / p.splits, sp = blockIdx.z - n * p.splits;
    const int o0 = blockIdx.x * WM, c0 = blockIdx.y * WN;
    const int P = p.ho * p.wo;
    const int pbeg = sp * p.chunk, pend = min(P, pbeg + p.chunk);
    const int pk = tid >> 4, q4 = (tid & 15) * 4;
    const int ty = tid >> 4, tx = tid & 15;

    long long coloff[4];
    bool colvalid[4];
#pragma unroll
    for (int j = 0; j < 4; j++) {
        int col = c0 + q4 + j;
        colvalid[j] = col < p.Kf;
        int tap = col / p.ci, c = col - tap * p.ci;
        int ky = tap / p.k, kx = tap - ky * p.k;
        coloff[j] = ((long long)ky * p.wp + kx) * p.ci + c;
    }
    float2 acc2[4][2];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) acc2[i][j] = make_float2(0.f, 0.f);

    for (int p0 = pbeg; p0 < pend; p0 += WK) {
        const int pix = p0 + pk;
        const bool pvalid = pix < pend;
        int oy = 0, ox = 0;
        if (pvalid) { oy = pix / p.wo; ox = pix - oy * p.wo; }
        const long long dbase = (((long long)n * p.dhp + p.dorg + oy) * p.dwp + p.dorg + ox) * p.co;
        const long long xbase = (((long long)n * p.hp + p.org + oy * p.stride) * p.wp + p.org + ox * p.stride) * p.ci;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int o = o0 + q4 + j;
            float v = 0.f;
            if (pvalid && o < p.co) {
                if (DFMT == SKIT_FMT_F32) v = __ldg(p.d0 + dbase + o);
                else v = __bfloat162float(p.dh[dbase + o]) + __bfloat162float(p.dl[dbase + o]);
            }
            As[pk][q4 + j] = v;
            float u = 0.f;
            if (pvalid && colvalid[j]) {
                if (XFMT == SKIT_FMT_F32) u = __ldg(p.x0 + xbase + coloff[j]);
                else u = __bfloat162float(p.xh[xbase + coloff[j]]) + __bfloat162float(p.xl[xbase + coloff[j]]);
            }
            Bs[pk][q4 + j] = u;
        }
        __syncthreads();
#pragma unroll
        for (int k2 = 0; k2 < WK; k2++) {
            const float4 a = *reinterpret_cast<const float4*>(&As[k2][ty * 4]);
            const float4 b = *reinterpret_cast<const float4*>(&Bs[k2][tx * 4]);
            const float av[4] = {a.x, a.y, a.z, a.w};
            const float2 bp[2] = {make_float2(b.x, b.y), make_float2(b.z, b.w)};      // column pairs: packed fp32 FMAs
#pragma unroll
            for (int i = 0; i < 4; i++) {
                const float2 aa = make_float2(av[i], av[i]);
#pragma unroll
                for (int j = 0; j < 2; j++) acc2[i][j] = ffma2(aa, bp[j], acc2[i][j]);
            }
        }
        __syncthreads();
    }
    float acc[4][4];
#pragma unroll
    for (int i = 0; i < 4; i++)
#pragma unroll
        for (int j = 0; j < 2; j++) { acc[i][2 * j] = acc2[i][j].x; acc[i][2 * j + 1] = acc2[i][j].y; }
#pragma unroll
    for (int i = 0; i < 4; i++) {
        int o = o0 + ty * 4 + i;
        if (o >= p.co) continue;
#pragma unroll
        for (int j = 0; j < 4; j++) {
            int col = c0 + tx * 4 + j;
            if (col >= p.Kf) continue;
            atomicAdd(p.dwf + (long long)col * p.co + o, acc[i][j]);
        }
    }
}

// Weight gradient of a layer with very few output channels (PatchGAN head 512 -> 1, generator head on the fp32 path):
// dwf[(tap*ci + c)*co + o] += sum_pix dy[pix][o] * x[pix + tap][c].  The 64x64 GEMM tile of wgrad_simt_kernel would carry
// one useful row; here a thread owns one input channel of one tap (coalesced 128-channel runs of x), keeps co (<= 4)
// accumulators, and walks a slice of the pixels.  grid: (taps, channel chunks of 128, n * pixel splits).
template <int XFMT, int DFMT, int CO>
__global__ void __launch_bounds__(128) wgrad_thin_co_kernel(WgradP p) {
    const int tap = blockIdx.x, c = blockIdx.y * 128 + threadIdx.x;
    const int n = blockIdx.z / p.splits, sp = blockIdx.z - n * p.splits;
    const int P = p.ho * p.wo;
    const int pbeg = sp * p.chunk, pend = min(P, pbeg + p.chunk);
    const int ky = tap / p.k, kx = tap - ky * p.k;
    if (c >= p.ci) return;
    float acc[CO];
#pragma unroll
    for (int o = 0; o < CO; o++) acc[o] = 0.f;
    for (int pix = pbeg; pix < pend; pix++) {
        const int oy = pix / p.wo, ox = pix - oy * p.wo;
        const long long dbase = (((long long)n * p.dhp + p.dorg + oy) * p.dwp + p.dorg + ox) * p.co;
        const long long xa = (((long long)n * p.hp + p.org + oy * p.stride + ky) * p.wp + p.org + ox * p.stride + kx) * p.ci + c;
        const float xv = (XFMT == SKIT_FMT_F32) ? __ldg(p.x0 + xa) : (__bfloat162float(p.xh[xa]) + __bfloat162float(p.xl[xa]));
#pragma unroll
        for (int o = 0; o < CO; o++) {
            if (o < p.co) {
                const float dv = (DFMT == SKIT_FMT_F32) ? __ldg(p.d0 + dbase + o) : (__bfloat162float(p.dh[dbase + o]) + __bfloat162float(p.dl[dbase + o]));
                acc[o] = fmaf(dv, xv, acc[o]);
            }
        }
    }
#pragma unroll
    for (int o = 0; o < CO; o++)
        if (o < p.co) atomicAdd(p.dwf + ((long long)tap * p.ci + c) * p.co + o, acc[o]);
}

// The same for ONE output channel, stride 1 and k*k <= 16 (the PatchGAN heads, 512 -> 1, k 4): the kernel above reads the whole
// activation once per tap (16 x 19.7 MB for the 98 x 98 head of the 768^2 step: 128 us).  Here a thread owns one input channel and
// walks INPUT positions: x[iy][ix][c] is loaded once and feeds all k*k taps, dw[ky][kx][c] += x[iy][ix][c] * dy[iy-ky][ix-kx], with
// the k x k window of dy (block-uniform, broadcast loads) sliding along the row in registers.  grid: (channel chunks of 128, n * row splits).
template <int XFMT, int DFMT, int K>
__global__ void __launch_bounds__(128) wgrad_thin_co1_s1_kernel(WgradP p, int rows_per_block) {
    const int c = blockIdx.x * 128 + threadIdx.x;
    const int hi = p.ho + K - 1, wi = p.wo + K - 1;
    const int nsplit = cdiv(hi, rows_per_block);
    const int n = blockIdx.y / nsplit, sp = blockIdx.y - n * nsplit;
    if (c >= p.ci) return;
    float acc[K][K];
#pragma unroll
    for (int a = 0; a < K; a++)
#pragma unroll
        for (int b2 = 0; b2 < K; b2++) acc[a][b2] = 0.f;
    const int iy_end = min(hi, (sp + 1) * rows_per_block);
    for (int iy = sp * rows_per_block; iy < iy_end; iy++) {
        float win[K][K];         // win[ky][kx] = dy[iy - ky][ix - kx] (0 outside the map)
#pragma unroll
        for (int a = 0; a < K; a++)
#pragma unroll
            for (int b2 = 0; b2 < K; b2++) win[a][b2] = 0.f;
        const long long xrow = (((long long)n * p.hp + p.org + iy) * p.wp + p.org) * p.ci + c;
        for (int ix = 0; ix < wi; ix++) {
#pragma unroll
            for (int a = 0; a < K; a++) {
#pragma unroll
                for (int b2 = K - 1; b2 > 0; b2--) win[a][b2] = win[a][b2 - 1];
                const int oy = iy - a;
                float dv = 0.f;
                if (oy >= 0 && oy < p.ho && ix < p.wo) {
                    const long long da = ((long long)n * p.dhp + p.dorg + oy) * p.dwp + p.dorg + ix;      // co == 1
                    dv = (DFMT == SKIT_FMT_F32) ? __ldg(p.d0 + da) : (__bfloat162float(p.dh[da]) + __bfloat162float(p.dl[da]));
                }
                win[a][0] = dv;
            }
            const long long xa = xrow + (long long)ix * p.ci;
            const float xv = (XFMT == SKIT_FMT_F32) ? __ldg(p.x0 + xa) : (__bfloat162float(p.xh[xa]) + __bfloat162float(p.xl[xa]));
#pragma unroll
            for (int a = 0; a < K; a++)
#pragma unroll
                for (int b2 = 0; b2 < K; b2++) acc[a][b2] = fmaf(win[a][b2], xv, acc[a][b2]);
        }
    }
#pragma unroll
    for (int a = 0; a < K; a++)
#pragma unroll
        for (int b2 = 0; b2 < K; b2++) atomicAdd(p.dwf + (long long)(a * K + b2) * p.ci + c, acc[a][b2]);
}

// dbias[o] += sum over pixels of dy (operand with halo).  grid: (pixel chunks, n), block 256 = 8 pixel
// lanes x 32 channel lanes: a warp reads 32 consecutive channels of one pixel (coalesced), the 8 pixel
// lanes are combined through shared memory before one atomic per channel per CTA.
template <int DFMT>
__global__ void __launch_bounds__(256) dbias_kernel(const float* d0, const __nv_bfloat16* dh, const __nv_bfloat16* dl,
                                                    int dhp, int dwp, int co, int cs, int dorg, int ho, int wo,
                                                    int chunk, float* dbias) {   // co channels summed, cs = channel stride
    __shared__ float red[8][33];
    const int n = blockIdx.y;
    const int P = ho * wo;
    const int pbeg = blockIdx.x * chunk, pend = min(P, pbeg + chunk);
    const int cl = threadIdx.x & 31, pl = threadIdx.x >> 5;
    for (int o0 = 0; o0 < co; o0 += 32) {
        const int o = o0 + cl;
        float s = 0.f;
        if (o < co) {
            for (int pix = pbeg + pl; pix < pend; pix += 8) {
                int oy = pix / wo, ox = pix - oy * wo;
                long long a = (((long long)n * dhp + dorg + oy) * dwp + dorg + ox) * cs + o;
                s += (DFMT == SKIT_FMT_F32) ? d0[a] : (__bfloat162float(dh[a]) + __bfloat162float(dl[a]));
            }
        }
        red[pl][cl] = s;
        __syncthreads();
        if (pl == 0 && o < co) {
            float t = 0.f;
#pragma unroll
            for (int i = 0; i < 8; i++) t += red[i][cl];
            atomicAdd(dbias + o, t);
        }
        __syncthreads();
    }
}

// Vector variant for channel counts that are multiples of 4: a thread owns 4 consecutive channels (one 16-byte fp32 load, or two
// 8-byte bf16 loads), 256 / (co / 4) pixel lanes per block; the pixel lanes are combined through shared memory.
template <int DFMT>
__global__ void __launch_bounds__(256) dbias_vec4_kernel(const float* __restrict__ d0, const __nv_bfloat16* __restrict__ dh,
                                                         const __nv_bfloat16* __restrict__ dl, int dhp, int dwp, int co, int cs, int dorg,
                                                         int ho, int wo, int chunk, float* dbias) {
    __shared__ float red[256 * 4];
    const int n = blockIdx.y;
    const int P = ho * wo;
    const int cv = co / 4;                 // <= 256
    const int PL = 256 / cv;
    const int cl = threadIdx.x % cv, pl = threadIdx.x / cv;
    const int pbeg = blockIdx.x * chunk, pend = min(P, pbeg + chunk);
    float s[4] = {0.f, 0.f, 0.f, 0.f};
    if (pl < PL) {
        for (int pix = pbeg + pl; pix < pend; pix += PL) {
            const int oy = pix / wo, ox = pix - oy * wo;
            const long long a = (((long long)n * dhp + dorg + oy) * dwp + dorg + ox) * cs + cl * 4;
            if (DFMT == SKIT_FMT_F32) {
                const float4 v = *reinterpret_cast<const float4*>(d0 + a);
                s[0] += v.x; s[1] += v.y; s[2] += v.z; s[3] += v.w;
            } else {
                const uint2 h = *reinterpret_cast<const uint2*>(dh + a), l = *reinterpret_cast<const uint2*>(dl + a);
                const __nv_bfloat162 h0 = *reinterpret_cast<const __nv_bfloat162*>(&h.x), h1 = *reinterpret_cast<const __nv_bfloat162*>(&h.y);
                const __nv_bfloat162 l0 = *reinterpret_cast<const __nv_bfloat162*>(&l.x), l1 = *reinterpret_cast<const __nv_bfloat162*>(&l.y);
                s[0] += __low2float(h0) + __low2float(l0); s[1] += __high2float(h0) + __high2float(l0);
                s[2] += __low2float(h1) + __low2float(l1); s[3] += __high2float(h1) + __high2float(l1);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < 4; j++) red[threadIdx.x * 4 + j] = s[j];
    __syncthreads();
    for (int t = threadIdx.x; t < co; t += 256) {
        float tot = 0.f;
        for (int q = 0; q < PL; q++) tot += red[(q * cv + t / 4) * 4 + (t & 3)];
        atomicAdd(dbias + t, tot);
    }
}

// ---------------------------------------------------------------------------------- weight packs
// Iterates in OUTPUT order (coalesced bf16 / fp32 stores; the scattered side is the L2-resident source).
// The bf16 packs are [tap][N][K] (K innermost), the fp32 packs [tap][K][N] (N innermost):
//   mode 0: K = conv ci, N = conv co.   modes 1-3 (input-gradient packs): K = conv co, N = conv ci.
__device__ __forceinline__ long long pack_src_index(int mode, int k, int co, int ci, int tap, int o, int c) {
    int ky, kx;
    if (mode == 0 || mode == 2) { ky = tap / k; kx = tap - ky * k; }
    else if (mode == 1) { const int f = k * k - 1 - tap; ky = f / k; kx = f - ky * k; }
    else {  // mode 3: tap = phase*(k/2)^2 + (k/2-1-ky/2)*(k/2) + (k/2-1-kx/2), phase = (ky%2)*2 + kx%2
        const int kh = k / 2, phase = tap / (kh * kh), r = tap - phase * kh * kh;
        const int a = r / kh, b2 = r - a * kh;
        ky = 2 * (kh - 1 - a) + (phase >> 1); kx = 2 * (kh - 1 - b2) + (phase & 1);
    }
    return (((long long)o * ci + c) * k + ky) * k + kx;
}

// x-folded packs (modes 4 / 5, bf16 only): [ky][N][64], K index j = kx*cp + ch
__device__ __forceinline__ float folded_pack_value(const float* __restrict__ w, int mode, int k, int co, int ci, int cp, int ky, int nn, int j) {
    const int kx = j / cp, ch = j - kx * cp;
    if (kx >= k) return 0.f;
    if (mode == 4) { return ch < ci ? __ldg(w + (((long long)nn * ci + ch) * k + ky) * k + kx) : 0.f; }
    return ch < co ? __ldg(w + (((long long)ch * ci + nn) * k + (k - 1 - ky)) * k + (k - 1 - kx)) : 0.f;
}

__global__ void pack_weights_folded_kernel(const float* __restrict__ w, int co, int ci, int k, int mode, int cp,
                                           __nv_bfloat16* hi, __nv_bfloat16* lo) {
    const int Nd = mode == 4 ? co : ci;
    const long long total = (long long)k * Nd * 64;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int j = (int)(i % 64); long long t = i / 64;
        const int nn = (int)(t % Nd); const int ky = (int)(t / Nd);
        __nv_bfloat16 h, l;
        split_bf16(folded_pack_value(w, mode, k, co, ci, cp, ky, nn, j), h, l);
        hi[i] = h; lo[i] = l;
    }
}

__global__ void pack_weights_kernel(const float* __restrict__ w, int co, int ci, int k, int mode,
                                    float* f32, __nv_bfloat16* hi, __nv_bfloat16* lo, int kpad) {
    const int Kd = mode == 0 ? ci : co, Nd = mode == 0 ? co : ci;
    const int Kp = kpad > Kd ? kpad : Kd;     // bf16 packs may pad the reduction axis with zeros (TMA row alignment)
    const long long total = (long long)k * k * Nd * Kp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        if (hi) {   // [tap][N][Kp]
            const int kk = (int)(i % Kp); long long t = i / Kp;
            const int nn = (int)(t % Nd); const int tap = (int)(t / Nd);
            const int o = mode == 0 ? nn : kk, c = mode == 0 ? kk : nn;
            const float v = kk < Kd ? __ldg(w + pack_src_index(mode, k, co, ci, tap, o, c)) : 0.f;
            __nv_bfloat16 h, l;
            split_bf16(v, h, l);
            hi[i] = h; lo[i] = l;
        }
        if (f32 && i < (long long)k * k * Nd * Kd) {  // [tap][K][N]
            const int nn = (int)(i % Nd); long long t = i / Nd;
            const int kk = (int)(t % Kd); const int tap = (int)(t / Kd);
            const int o = mode == 0 ? nn : kk, c = mode == 0 ? kk : nn;
            f32[i] = __ldg(w + pack_src_index(mode, k, co, ci, tap, o, c));
        }
    }
}

// One launch for every pack of a net: `descs` (device) lists the packs with their prefix offsets into a single index space.
// A block owns a contiguous chunk of that space: one descriptor search per chunk, then it only steps forward.
constexpr int kPackChunk = 4096;
__global__ void __launch_bounds__(256) pack_weights_batched_kernel(const skit_pack_desc* __restrict__ descs, int n, long long total) {
    __shared__ int s_first;
    for (long long base = (long long)blockIdx.x * kPackChunk; base < total; base += (long long)gridDim.x * kPackChunk) {
        if (threadIdx.x == 0) {
            int lo = 0, hi = n - 1;
            while (lo < hi) {   // last descriptor whose start <= base
                const int mid = (lo + hi + 1) >> 1;
                if (descs[mid].start <= base) lo = mid; else hi = mid - 1;
            }
            s_first = lo;
        }
        __syncthreads();
        int di = s_first;
        const long long end = min(total, base + kPackChunk);
        for (long long i = base + threadIdx.x; i < end; i += 256) {
            while (di + 1 < n && descs[di + 1].start <= i) di++;
            const skit_pack_desc& d = descs[di];
            const long long j = i - d.start;
            const int mode = d.mode, k = d.k, co = d.co, ci = d.ci;
            if (mode >= 4) {   // x-folded bf16 pack [ky][N][64]; d.reserved = channels per folded column
                const int Nd4 = mode == 4 ? co : ci;
                const int jj = (int)(j % 64); long long t4 = j / 64;
                const int nn4 = (int)(t4 % Nd4); const int ky4 = (int)(t4 / Nd4);
                __nv_bfloat16 h4, l4;
                split_bf16(folded_pack_value(d.w, mode, k, co, ci, d.reserved, ky4, nn4, jj), h4, l4);
                ((__nv_bfloat16*)d.hi)[j] = h4; ((__nv_bfloat16*)d.lo)[j] = l4;
                continue;
            }
            const int Kd = mode == 0 ? ci : co, Nd = mode == 0 ? co : ci;
            if (d.hi) {   // bf16 hi/lo [tap][N][Kp]
                const int Kp = d.kpad > Kd ? d.kpad : Kd;
                const int kk = (int)(j % Kp); long long t = j / Kp;
                const int nn = (int)(t % Nd); const int tap = (int)(t / Nd);
                const int o = mode == 0 ? nn : kk, c = mode == 0 ? kk : nn;
                const float v = kk < Kd ? __ldg(d.w + pack_src_index(mode, k, co, ci, tap, o, c)) : 0.f;
                __nv_bfloat16 h, l;
                split_bf16(v, h, l);
                ((__nv_bfloat16*)d.hi)[j] = h; ((__nv_bfloat16*)d.lo)[j] = l;
            } else {      // fp32 [tap][K][N]
                const int nn = (int)(j % Nd); long long t = j / Nd;
                const int kk = (int)(t % Kd); const int tap = (int)(t / Kd);
                const int o = mode == 0 ? nn : kk, c = mode == 0 ? kk : nn;
                d.f32[j] = __ldg(d.w + pack_src_index(mode, k, co, ci, tap, o, c));
            }
        }
        __syncthreads();
    }
}

// Tiled refresh for k*k <= 16 packs (modes 0-3): a block stages a [32 o][32 c][k*k] brick of the reference-layout weight in
// shared memory with coalesced loads (a brick row is 32*k*k contiguous floats) and writes every tap's 32x32 slab of the
// pack from there — the element-wise kernel above reads the source at a k*k-float stride, one useful float per sector.
// descs[i].start = prefix sum of TILES (ceil(N/32) * ceil(Kpadded/32) per pack).
constexpr int kPackTileRow = 33, kPackTilePlane = 32 * kPackTileRow + 4;
__device__ __forceinline__ int pack_dst_tap(int mode, int k, int t) {
    if (mode == 0 || mode == 2) return t;
    if (mode == 1) return k * k - 1 - t;
    const int ky = t / k, kx = t - ky * k, kh = k / 2;
    return ((ky & 1) * 2 + (kx & 1)) * kh * kh + (kh - 1 - ky / 2) * kh + (kh - 1 - kx / 2);
}
__global__ void __launch_bounds__(256) pack_weights_tiled_kernel(const skit_pack_desc* __restrict__ descs, int n, int total_tiles) {
    extern __shared__ float tile[];   // [k*k][32 o][33] (+4 per plane)
    __shared__ int s_desc;
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    for (int tix = blockIdx.x; tix < total_tiles; tix += gridDim.x) {
        if (tid == 0) {
            int lo = 0, hi = n - 1;
            while (lo < hi) {
                const int mid = (lo + hi + 1) >> 1;
                if (descs[mid].start <= tix) lo = mid; else hi = mid - 1;
            }
            s_desc = lo;
        }
        __syncthreads();
        const skit_pack_desc& d = descs[s_desc];
        const int mode = d.mode, k = d.k, co = d.co, ci = d.ci, kk = k * k;
        const int Kd = mode == 0 ? ci : co, Nd = mode == 0 ? co : ci;
        const int Kp = (d.hi && d.kpad > Kd) ? d.kpad : Kd;
        const int tilesK = (Kp + 31) >> 5;
        const int local = tix - (int)d.start;
        const int tn = local / tilesK, tk = local - tn * tilesK;
        const int o0 = (mode == 0 ? tn : tk) * 32, c0 = (mode == 0 ? tk : tn) * 32;
        // stage: rows i (o), each 32*kk contiguous source floats
        const int rowlen = 32 * kk;
        for (int i = warp; i < 32; i += 8) {
            const int o = o0 + i;
            const float* src = d.w + ((long long)o * ci + c0) * kk;
            const int valid = o < co ? min(32, ci - c0) * kk : 0;
            for (int e = lane; e < rowlen; e += 32) {
                const int j = e / kk, t = e - j * kk;
                tile[t * kPackTilePlane + i * kPackTileRow + j] = e < valid ? __ldg(src + e) : 0.f;
            }
        }
        __syncthreads();
        const bool lanes_along_c = (mode == 0);   // bf16 pack: K innermost (c for mode 0, o otherwise)
        if (d.hi) {
            __nv_bfloat16* hi = (__nv_bfloat16*)d.hi; __nv_bfloat16* lo = (__nv_bfloat16*)d.lo;
            // a warp writes two N rows per pass: 16 lanes x bf16x2 each
            const int q = lane & 15, half = lane >> 4;
            for (int pr = warp; pr < kk * 16; pr += 8) {
                const int t = pr >> 4, r = (pr & 15) * 2 + half;     // r: index along N within the tile
                const int nn = (mode == 0 ? o0 : c0) + r, kbase = (mode == 0 ? c0 : o0) + 2 * q;
                if (nn >= Nd || kbase >= Kp) continue;
                float v0, v1;
                if (lanes_along_c) { v0 = tile[t * kPackTilePlane + r * kPackTileRow + 2 * q]; v1 = tile[t * kPackTilePlane + r * kPackTileRow + 2 * q + 1]; }
                else { v0 = tile[t * kPackTilePlane + (2 * q) * kPackTileRow + r]; v1 = tile[t * kPackTilePlane + (2 * q + 1) * kPackTileRow + r]; }
                __nv_bfloat16 h0, l0, h1, l1;
                split_bf16(v0, h0, l0); split_bf16(v1, h1, l1);
                const long long off = ((long long)pack_dst_tap(mode, k, t) * Nd + nn) * Kp + kbase;
                if (kbase + 1 < Kp) {
                    *reinterpret_cast<__nv_bfloat162*>(hi + off) = __halves2bfloat162(h0, h1);
                    *reinterpret_cast<__nv_bfloat162*>(lo + off) = __halves2bfloat162(l0, l1);
                } else { hi[off] = h0; lo[off] = l0; }
            }
        }
        if (d.f32 && (!d.hi || tk * 32 < Kd)) {   // fp32 pack [tap][K][N]: N innermost (o for mode 0, c otherwise)
            for (int pr = warp; pr < kk * 32; pr += 8) {
                const int t = pr >> 5, r = pr & 31;                   // r: index along K within the tile
                const int kidx = (mode == 0 ? c0 : o0) + r, nn = (mode == 0 ? o0 : c0) + lane;
                if (kidx >= Kd || nn >= Nd) continue;
                const float v = mode == 0 ? tile[t * kPackTilePlane + lane * kPackTileRow + r] : tile[t * kPackTilePlane + r * kPackTileRow + lane];
                d.f32[((long long)pack_dst_tap(mode, k, t) * Kd + kidx) * Nd + nn] = v;
            }
        }
        __syncthreads();
    }
}

// layout 0: dwf[(tap*cip + c)*cop + o]   layout 1: dwf[(tap*cop + o)*cip + c]   (cop/cip: channel counts of the
// operands the partial sums were computed on, >= the real co/ci when those were zero-padded)
__global__ void unpack_wgrad_kernel(const float* __restrict__ dwf, int co, int ci, int k, float* dw, int accumulate, int layout,
                                    int cop, int cip) {
    const long long total = (long long)co * ci * k * k;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int kx = i % k; long long t = i / k;
        int ky = t % k; t /= k;
        int c = t % ci; int o = t / ci;
        const long long tap = ky * k + kx;
        float v = layout == 0 ? dwf[(tap * cip + c) * cop + o] : dwf[(tap * cop + o) * cip + c];
        if (accumulate) atomicAdd(dw + i, v);   // passes running on parallel streams accumulate into the same bucket
        else dw[i] = v;
    }
}

// Tiled variant: a block owns 32 output channels x TC input channels x all taps, reads the partial sums along their contiguous
// axis (c for layout 1, o for layout 0) into shared memory and writes each output channel's TC * k * k consecutive weights as one
// run — both sides coalesced (the element-wise kernel reads with a stride of cip * cop floats between neighbouring threads).
template <int TC>
__global__ void __launch_bounds__(256) unpack_wgrad_tiled_kernel(const float* __restrict__ dwf, int co, int ci, int kk, float* dw, int accumulate,
                                                                 int layout, int cop, int cip) {
    extern __shared__ float tile[];      // [32 o][TC c][kk | 1]
    const int kp = kk | 1;
    const int o0 = blockIdx.y * 32, c0 = blockIdx.x * TC;
    const int n = 32 * TC * kk;
    for (int i = threadIdx.x; i < n; i += 256) {
        int tap, ol, cl;
        if (layout == 1) { cl = i % TC; const int t = i / TC; ol = t % 32; tap = t / 32; }
        else { ol = i % 32; const int t = i / 32; cl = t % TC; tap = t / TC; }
        const int o = o0 + ol, c = c0 + cl;
        float v = 0.f;
        if (o < co && c < ci) v = layout == 0 ? dwf[((long long)tap * cip + c) * cop + o] : dwf[((long long)tap * cop + o) * cip + c];
        tile[(ol * TC + cl) * kp + tap] = v;
    }
    __syncthreads();
    const int run = TC * kk;
    for (int i = threadIdx.x; i < 32 * run; i += 256) {
        const int ol = i / run, j = i - ol * run;
        const int cl = j / kk, tap = j - cl * kk;
        const int o = o0 + ol, c = c0 + cl;
        if (o >= co || c >= ci) continue;
        const float v = tile[(ol * TC + cl) * kp + tap];
        float* dst = dw + ((long long)o * ci + c) * kk + tap;
        if (accumulate) atomicAdd(dst, v);
        else *dst = v;
    }
}

// dwf from wgrad_tc on a folded operand: taps = ky, x channels j = kx*cp + c.  layout 0: [(ky*64 + j)*cop + o], 1: [(ky*cop + o)*64 + j]
__global__ void unpack_wgrad_folded_kernel(const float* __restrict__ dwf, int co, int ci, int k, int kw, int cp, float* dw, int layout, int cop) {
    const long long total = (long long)co * ci * k * kw;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int kx = i % kw; long long t = i / kw;
        int ky = t % k; t /= k;
        int c = t % ci; int o = t / ci;
        const int j = kx * cp + c;
        const float v = layout == 0 ? dwf[((long long)ky * 64 + j) * cop + o] : dwf[((long long)ky * cop + o) * 64 + j];
        atomicAdd(dw + i, v);
    }
}

// dwf from wgrad_tc with an x-folded GRADIENT operand (taps = ky; gradient channels j = (kw-1-kx)*cp + o):
// layout 0: [(ky*cip + c)*64 + j], layout 1: [(ky*64 + j)*cip + c]
__global__ void unpack_wgrad_dyfolded_kernel(const float* __restrict__ dwf, int co, int ci, int k, int kw, int cp, float* dw, int layout, int cip) {
    const long long total = (long long)co * ci * k * kw;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        int kx = i % kw; long long t = i / kw;
        int ky = t % k; t /= k;
        int c = t % ci; int o = t / ci;
        const int j = (kw - 1 - kx) * cp + o;
        const float v = layout == 0 ? dwf[((long long)ky * cip + c) * 64 + j] : dwf[((long long)ky * 64 + j) * cip + c];
        atomicAdd(dw + i, v);
    }
}

int conv_fwd_simt(const skit_operand* x, const skit_weights* w, int stride, int org, int ho, int wo,
                  const float* bias, float* y, double* stats, int stats_mode, cudaStream_t st) {
    ConvP p{};
    p.x0 = (const float*)x->p0; p.xh = (const __nv_bfloat16*)x->p0; p.xl = (const __nv_bfloat16*)x->p1;
    p.hp = x->hp; p.wp = x->wp; p.ci = x->c;
    p.w = w->f32; p.bias = bias; p.y = y; p.stats = stats; p.stats_per_n = stats_mode == SKIT_NORM_INSTANCE;
    p.k = w->k; p.stride = stride; p.org = org; p.ho = ho; p.wo = wo;
    p.ncol = w->co; p.K = w->k * w->k * w->ci; p.M = ho * wo; p.arows = 0;
    if (p.ncol <= 8 && p.ci % 32 == 0 && !stats) {
        if (x->fmt == SKIT_FMT_F32) return launch_thin<0, SKIT_FMT_F32>(p, x->n, st);
        return launch_thin<0, SKIT_FMT_BF16X2>(p, x->n, st);
    }
    p.n_img = x->n;
    p.flat = (x->n > 1 && p.M < 4 * BM && !p.stats_per_n) ? 1 : 0;  // many small images: tile over the whole batch
    const int bn = narrow_tile(p.ncol) ? 16 : BN;
    dim3 grid(p.flat ? cdiv(x->n * p.M, BM) : cdiv(p.M, BM), cdiv(p.ncol, bn), p.flat ? 1 : x->n);
    if (x->fmt == SKIT_FMT_F32) {
        if (bn == 16) conv_simt_kernel<0, SKIT_FMT_F32, 1><<<grid, 256, 0, st>>>(p);
        else conv_simt_kernel<0, SKIT_FMT_F32><<<grid, 256, 0, st>>>(p);
    } else conv_simt_kernel<0, SKIT_FMT_BF16X2><<<grid, 256, 0, st>>>(p);
    return check_launch("conv_simt_kernel<fwd>");
}

}  // namespace skit

using namespace skit;

extern "C" int skit_pack_conv_weights(const float* w, int co, int ci, int k, int mode,
                                      float* f32, void* hi, void* lo, void* stream) {
    SKIT_REQUIRE(w && co > 0 && ci > 0 && k > 0 && mode >= 0 && mode <= 3 && (mode != 3 || k % 2 == 0), "pack_conv_weights: bad arguments");
    SKIT_REQUIRE((hi == nullptr) == (lo == nullptr), "pack_conv_weights: hi and lo must be given together");
    long long total = (long long)co * ci * k * k;
    int blocks = (int)min((long long)148 * 8, cdivll(total, 256));
    pack_weights_kernel<<<blocks, 256, 0, as_stream(stream)>>>(w, co, ci, k, mode, f32, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, 0);
    return check_launch("pack_weights_kernel");
}

extern "C" int skit_pack_conv_weights_batched(const skit_pack_desc* descs_dev, int n, long long total, void* stream) {
    SKIT_REQUIRE(descs_dev && n > 0 && total > 0, "pack_conv_weights_batched: bad arguments");
    int blocks = (int)min((long long)148 * 8, cdivll(total, kPackChunk));
    pack_weights_batched_kernel<<<blocks, 256, 0, as_stream(stream)>>>(descs_dev, n, total);
    return check_launch("pack_weights_batched_kernel");
}

extern "C" int skit_pack_conv_weights_tiled(const skit_pack_desc* descs_dev, int n, int total_tiles, int max_k, void* stream) {
    SKIT_REQUIRE(descs_dev && n > 0 && total_tiles > 0 && max_k > 0 && max_k <= 4, "pack_conv_weights_tiled: bad arguments");
    static bool attr_set = false;
    const int smem = max_k * max_k * kPackTilePlane * (int)sizeof(float);
    if (!attr_set) {
        cudaError_t e = cudaFuncSetAttribute(pack_weights_tiled_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 16 * kPackTilePlane * (int)sizeof(float));
        if (e != cudaSuccess) {
            set_error("cudaFuncSetAttribute(pack_weights_tiled_kernel) failed: %s", cudaGetErrorString(e));
            return SKIT_ERR_CUDA;
        }
        attr_set = true;
    }
    const int blocks = min(total_tiles, 148 * 4);
    pack_weights_tiled_kernel<<<blocks, 256, smem, as_stream(stream)>>>(descs_dev, n, total_tiles);
    return check_launch("pack_weights_tiled_kernel");
}

extern "C" int skit_pack_conv_weights_folded(const float* w, int co, int ci, int k, int mode, int cp, void* hi, void* lo, void* stream) {
    SKIT_REQUIRE(w && hi && lo && co > 0 && ci > 0 && k > 0 && (mode == 4 || mode == 5), "pack_conv_weights_folded: bad arguments");
    SKIT_REQUIRE(cp >= (mode == 4 ? ci : co) && k * cp <= 64, "pack_conv_weights_folded: cp %d must hold the thin side and k*cp <= 64", cp);
    long long total = (long long)k * (mode == 4 ? co : ci) * 64;
    int blocks = (int)min((long long)148 * 8, cdivll(total, 256));
    pack_weights_folded_kernel<<<blocks, 256, 0, as_stream(stream)>>>(w, co, ci, k, mode, cp, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo);
    return check_launch("pack_weights_folded_kernel");
}

extern "C" int skit_pack_conv_weights_padded(const float* w, int co, int ci, int k, int mode, int kpad, void* hi, void* lo, void* stream) {
    SKIT_REQUIRE(w && hi && lo && co > 0 && ci > 0 && k > 0 && mode >= 0 && mode <= 3 && (mode != 3 || k % 2 == 0), "pack_conv_weights_padded: bad arguments");
    const int Kd = mode == 0 ? ci : co, Nd = mode == 0 ? co : ci;
    SKIT_REQUIRE(kpad >= Kd && kpad % 8 == 0, "pack_conv_weights_padded: kpad %d must be a multiple of 8 and >= %d", kpad, Kd);
    long long total = (long long)k * k * Nd * kpad;
    int blocks = (int)min((long long)148 * 8, cdivll(total, 256));
    pack_weights_kernel<<<blocks, 256, 0, as_stream(stream)>>>(w, co, ci, k, mode, nullptr, (__nv_bfloat16*)hi, (__nv_bfloat16*)lo, kpad);
    return check_launch("pack_weights_kernel");
}

static int unpack_wgrad(const float* dwf, int co, int ci, int k, float* dw, int accumulate, int layout, cudaStream_t st,
                        int cop = 0, int cip = 0) {
    long long total = (long long)co * ci * k * k;
    const int kk = k * k;
    static const bool tiled = !(getenv("SKIT_UNPACK_TILED") && atoi(getenv("SKIT_UNPACK_TILED")) == 0);
    if (tiled && co >= 32 && ci >= 16 && kk <= 16) {      // the trunk / discriminator layers: both sides coalesced through a shared-memory tile
        dim3 grid(cdiv(ci, 16), cdiv(co, 32));
        unpack_wgrad_tiled_kernel<16><<<grid, 256, (size_t)32 * 16 * (kk | 1) * sizeof(float), st>>>(dwf, co, ci, kk, dw, accumulate, layout,
                                                                                                    cop ? cop : co, cip ? cip : ci);
        return check_launch("unpack_wgrad_tiled_kernel");
    }
    int blocks = (int)min((long long)148 * 8, cdivll(total, 256));
    unpack_wgrad_kernel<<<blocks, 256, 0, st>>>(dwf, co, ci, k, dw, accumulate, layout, cop ? cop : co, cip ? cip : ci);
    return check_launch("unpack_wgrad_kernel");
}

extern "C" int skit_unpack_conv_wgrad(const float* dwf, int co, int ci, int k, float* dw, int accumulate, void* stream) {
    SKIT_REQUIRE(dwf && dw && co > 0 && ci > 0 && k > 0, "unpack_conv_wgrad: bad arguments");
    return unpack_wgrad(dwf, co, ci, k, dw, accumulate, 0, as_stream(stream));
}

extern "C" int skit_conv2d_dgrad_gather(const float* dy, int n, int ho, int wo, int co,
                                        const skit_weights* wg, int stride, int hp, int wp, float* dx, void* stream) {
    SKIT_REQUIRE(dy && wg && wg->f32 && dx, "conv2d_dgrad_gather: null pointer");
    // mode-2 pack: the GEMM reduces over (tap, dy channel) and produces the conv's input channels
    SKIT_REQUIRE(wg->ci == co, "conv2d_dgrad_gather: weight pack reduces over %d channels but dy has %d", wg->ci, co);
    ConvP p{};
    p.x0 = dy; p.hp = hp; p.wp = wp; p.ci = wg->co;
    p.w = wg->f32; p.bias = nullptr; p.y = dx; p.stats = nullptr;
    p.k = wg->k; p.stride = stride; p.org = 0; p.ho = ho; p.wo = wo;
    p.ncol = wg->co; p.K = wg->k * wg->k * co; p.M = hp * wp; p.arows = co;
    if (p.ncol <= 8 && co % 32 == 0) return launch_thin<1, SKIT_FMT_F32>(p, n, as_stream(stream));
    p.n_img = n;
    if (stride > 1 && wg->k % stride == 0) {   // parity classes: 1/stride^2 of the taps per class
        p.phased = 1; p.flat = 0;
        const int mclass = cdiv(hp, stride) * cdiv(wp, stride);
        const int bn = narrow_tile(p.ncol) ? 16 : BN;
        dim3 grid(cdiv(mclass, BM), cdiv(p.ncol, bn), n * stride * stride);
        if (bn == 16) conv_simt_kernel<1, SKIT_FMT_F32, 1><<<grid, 256, 0, as_stream(stream)>>>(p);
        else conv_simt_kernel<1, SKIT_FMT_F32><<<grid, 256, 0, as_stream(stream)>>>(p);
        return check_launch("conv_simt_kernel<dgrad,phased>");
    }
    p.flat = (n > 1 && p.M < 4 * BM) ? 1 : 0;
    dim3 grid(p.flat ? cdiv(n * p.M, BM) : cdiv(p.M, BM), cdiv(p.ncol, BN), p.flat ? 1 : n);
    conv_simt_kernel<1, SKIT_FMT_F32><<<grid, 256, 0, as_stream(stream)>>>(p);
    return check_launch("conv_simt_kernel<dgrad>");
}

// ConvTranspose2d forward (thirdparty/unet/unet_parts_custom.py:63; F.conv_transpose2d k4 s2 p1 of the default
// U-Net generator) = the gather form of a strided conv's input gradient, cropped by `pad`, plus bias, with the
// InstanceNorm statistics of the result accumulated in the same epilogue.
extern "C" int skit_conv_transpose2d_fwd(const float* x, int n, int h, int w, int ci, const skit_weights* wg, int stride, int pad,
                                         int ho, int wo, const float* bias, float* y, int y_ctot, int y_c0,
                                         double* stats, int stats_mode, void* stream) {
    SKIT_REQUIRE(x && wg && wg->f32 && y && n > 0 && h > 0 && w > 0 && ci > 0, "conv_transpose2d_fwd: bad arguments");
    SKIT_REQUIRE(wg->ci == ci, "conv_transpose2d_fwd: weight pack reduces over %d channels but x has %d", wg->ci, ci);
    SKIT_REQUIRE(ho == (h - 1) * stride - 2 * pad + wg->k && wo == (w - 1) * stride - 2 * pad + wg->k,
                 "conv_transpose2d_fwd: output %dx%d does not match (in-1)*stride - 2*pad + k", ho, wo);
    SKIT_REQUIRE(y_c0 >= 0 && y_c0 + wg->co <= y_ctot, "conv_transpose2d_fwd: output slice exceeds y_ctot");
    SKIT_REQUIRE(stats == nullptr || stats_mode == SKIT_NORM_INSTANCE || stats_mode == SKIT_NORM_BATCH, "conv_transpose2d_fwd: stats without mode");
    ConvP p{};
    p.x0 = x; p.hp = ho; p.wp = wo; p.ci = wg->co;
    p.w = wg->f32; p.bias = bias; p.y = y; p.stats = stats; p.stats_per_n = stats_mode == SKIT_NORM_INSTANCE;
    p.k = wg->k; p.stride = stride; p.org = 0; p.ho = h; p.wo = w;
    p.ncol = wg->co; p.K = wg->k * wg->k * ci; p.M = ho * wo; p.arows = ci;
    p.crop = pad; p.yc = y_ctot; p.yoff = y_c0;
    p.n_img = n; p.flat = 0;
    if (stride > 1 && wg->k % stride == 0) {
        p.phased = 1;
        const int mclass = cdiv(ho, stride) * cdiv(wo, stride);
        const int bn = narrow_tile(p.ncol) ? 16 : BN;
        dim3 grid(cdiv(mclass, BM), cdiv(p.ncol, bn), n * stride * stride);
        if (bn == 16) conv_simt_kernel<1, SKIT_FMT_F32, 1><<<grid, 256, 0, as_stream(stream)>>>(p);
        else conv_simt_kernel<1, SKIT_FMT_F32><<<grid, 256, 0, as_stream(stream)>>>(p);
        return check_launch("conv_simt_kernel<convT,phased>");
    }
    dim3 grid(cdiv(p.M, BM), cdiv(p.ncol, BN), n);
    conv_simt_kernel<1, SKIT_FMT_F32><<<grid, 256, 0, as_stream(stream)>>>(p);
    return check_launch("conv_simt_kernel<convT>");
}

namespace skit {
static int launch_dbias(const skit_operand* dy, int dy_org, int ho, int wo, int nch, float* dbias, cudaStream_t st) {
    const int P = ho * wo, n = dy->n;
    const int chunk = max(64, cdiv(P, max(1, (148 * 4) / n)));
    dim3 grid(cdiv(P, chunk), n);
    const bool f32 = dy->fmt == SKIT_FMT_F32;
    const float* d0 = f32 ? (const float*)dy->p0 : nullptr;
    const __nv_bfloat16* dh = f32 ? nullptr : (const __nv_bfloat16*)dy->p0;
    const __nv_bfloat16* dl = f32 ? nullptr : (const __nv_bfloat16*)dy->p1;
    if (nch % 4 == 0 && dy->c % 4 == 0 && nch <= 1024 && 256 % (nch / 4) == 0) {
        if (f32) dbias_vec4_kernel<0><<<grid, 256, 0, st>>>(d0, dh, dl, dy->hp, dy->wp, nch, dy->c, dy_org, ho, wo, chunk, dbias);
        else dbias_vec4_kernel<1><<<grid, 256, 0, st>>>(d0, dh, dl, dy->hp, dy->wp, nch, dy->c, dy_org, ho, wo, chunk, dbias);
        return check_launch("dbias_vec4_kernel");
    }
    if (f32) dbias_kernel<0><<<grid, 256, 0, st>>>(d0, dh, dl, dy->hp, dy->wp, nch, dy->c, dy_org, ho, wo, chunk, dbias);
    else dbias_kernel<1><<<grid, 256, 0, st>>>(d0, dh, dl, dy->hp, dy->wp, nch, dy->c, dy_org, ho, wo, chunk, dbias);
    return check_launch("dbias_kernel");
}
}  // namespace skit

// dbias[o] += sum over (n, pixels) of an operand's interior (bias gradient of layers without a norm).
extern "C" int skit_dbias(const skit_operand* dy, int dy_org, int ho, int wo, float* dbias, void* stream) {
    SKIT_REQUIRE(dy && dy->p0 && dbias && ho > 0 && wo > 0, "dbias: bad arguments");
    SKIT_REQUIRE(dy_org + ho <= dy->hp && dy_org + wo <= dy->wp, "dbias: window exceeds the operand");
    return launch_dbias(dy, dy_org, ho, wo, dy->c, dbias, as_stream(stream));
}

namespace skit {
int wgrad_tc(const skit_operand* x, int org, const skit_operand* dy, int dy_org, int k, int stride,
             int ho, int wo, float* partial, int* layout, cudaStream_t st, int kw = 0, int dy_org_y = -1);  // tc_wgrad.cu
bool wgrad_tc_eligible(const skit_operand* x, const skit_operand* dy, int k, int stride, int ho, int wo);
}

extern "C" int skit_conv2d_wgrad_ex(const skit_operand* x, int org, const skit_operand* dy, int dy_org,
                                    int k, int stride, int ho, int wo, float* scratch, float* dw, float* dbias, int impl,
                                    int co_real, int ci_real, void* stream);

extern "C" int skit_conv2d_wgrad(const skit_operand* x, int org, const skit_operand* dy, int dy_org,
                                 int k, int stride, int ho, int wo, float* scratch, float* dw, float* dbias, int impl, void* stream) {
    SKIT_REQUIRE(x && dy, "conv2d_wgrad: null pointer");
    return skit_conv2d_wgrad_ex(x, org, dy, dy_org, k, stride, ho, wo, scratch, dw, dbias, impl, dy->c, x->c, stream);
}

extern "C" int skit_conv2d_wgrad_ex(const skit_operand* x, int org, const skit_operand* dy, int dy_org,
                                    int k, int stride, int ho, int wo, float* scratch, float* dw, float* dbias, int impl,
                                    int co_real, int ci_real, void* stream) {
    SKIT_REQUIRE(x && dy && scratch && dw && x->p0 && dy->p0, "conv2d_wgrad: null pointer");
    SKIT_REQUIRE(co_real > 0 && co_real <= dy->c && ci_real > 0 && ci_real <= x->c, "conv2d_wgrad: real channel counts exceed the operands'");
    float* dwf = scratch;
    int layout = 0;
    SKIT_REQUIRE(x->n == dy->n, "conv2d_wgrad: batch mismatch");
    cudaStream_t st = as_stream(stream);
    const int n = x->n, co = dy->c, ci = x->c, P = ho * wo;
    bool tc = impl != SKIT_IMPL_SIMT && wgrad_tc_eligible(x, dy, k, stride, ho, wo);
    if (impl == SKIT_IMPL_TC && !tc) {
        set_error("conv2d_wgrad: shape not eligible for the tcgen05 path");
        return SKIT_ERR_UNSUPPORTED;
    }
    if (tc) {
        int rc = wgrad_tc(x, org, dy, dy_org, k, stride, ho, wo, dwf, &layout, st);
        if (rc) return rc;
    } else {
        WgradP p{};
        p.x0 = (const float*)x->p0; p.xh = (const __nv_bfloat16*)x->p0; p.xl = (const __nv_bfloat16*)x->p1;
        p.hp = x->hp; p.wp = x->wp; p.ci = ci; p.org = org;
        p.d0 = (const float*)dy->p0; p.dh = (const __nv_bfloat16*)dy->p0; p.dl = (const __nv_bfloat16*)dy->p1;
        p.dhp = dy->hp; p.dwp = dy->wp; p.co = co; p.dorg = dy_org;
        p.k = k; p.stride = stride; p.ho = ho; p.wo = wo; p.Kf = k * k * ci; p.dwf = dwf;
        static const bool co1_path = !(getenv("SKIT_WGRAD_CO1") && atoi(getenv("SKIT_WGRAD_CO1")) == 0);
        if (co == 1 && ci >= 64 && stride == 1 && (k == 4 || k == 3) && co1_path) {
            // one output channel, stride 1: every activation element read once for all taps
            const int hi = ho + k - 1;
            const int chunks = cdiv(ci, 128);
            const int want = max(1, (148 * 8) / max(1, chunks * n));
            const int rpb = max(1, cdiv(hi, want));
            dim3 grid(chunks, n * cdiv(hi, rpb));
            const bool xf = x->fmt == SKIT_FMT_F32, df = dy->fmt == SKIT_FMT_F32;
#define SKIT_CO1(KK)                                                                                         \
            do {                                                                                             \
                if (xf && df) wgrad_thin_co1_s1_kernel<0, 0, KK><<<grid, 128, 0, st>>>(p, rpb);              \
                else if (xf) wgrad_thin_co1_s1_kernel<0, 1, KK><<<grid, 128, 0, st>>>(p, rpb);               \
                else if (df) wgrad_thin_co1_s1_kernel<1, 0, KK><<<grid, 128, 0, st>>>(p, rpb);               \
                else wgrad_thin_co1_s1_kernel<1, 1, KK><<<grid, 128, 0, st>>>(p, rpb);                       \
            } while (0)
            if (k == 4) SKIT_CO1(4); else SKIT_CO1(3);
#undef SKIT_CO1
            int rc = check_launch("wgrad_thin_co1_s1_kernel");
            if (rc) return rc;
        } else if (co <= 4 && ci >= 64) {   // thin output side: channel-parallel reduction instead of a GEMM tile
            const int blocks = k * k * cdiv(ci, 128);
            int splits = max(1, min(cdiv(cdiv(148 * 8, blocks), n), cdiv(P, 32)));
            p.chunk = cdiv(P, splits);
            p.splits = cdiv(P, p.chunk);
            dim3 grid(k * k, cdiv(ci, 128), n * p.splits);
            if (x->fmt == SKIT_FMT_F32 && dy->fmt == SKIT_FMT_F32) wgrad_thin_co_kernel<0, 0, 4><<<grid, 128, 0, st>>>(p);
            else if (x->fmt == SKIT_FMT_F32) wgrad_thin_co_kernel<0, 1, 4><<<grid, 128, 0, st>>>(p);
            else if (dy->fmt == SKIT_FMT_F32) wgrad_thin_co_kernel<1, 0, 4><<<grid, 128, 0, st>>>(p);
            else wgrad_thin_co_kernel<1, 1, 4><<<grid, 128, 0, st>>>(p);
            int rc = check_launch("wgrad_thin_co_kernel");
            if (rc) return rc;
        } else {
        int tiles = cdiv(co, WM) * cdiv(p.Kf, WN);
        int want = cdiv(148 * 4, tiles);
        int splits = max(1, min(cdiv(want, n), cdiv(P, 64)));
        p.chunk = cdiv(cdiv(P, splits), WK) * WK;
        p.splits = cdiv(P, p.chunk);
        dim3 grid(cdiv(co, WM), cdiv(p.Kf, WN), n * p.splits);
        if (x->fmt == SKIT_FMT_F32 && dy->fmt == SKIT_FMT_F32) wgrad_simt_kernel<0, 0><<<grid, 256, 0, st>>>(p);
        else if (x->fmt == SKIT_FMT_F32) wgrad_simt_kernel<0, 1><<<grid, 256, 0, st>>>(p);
        else if (dy->fmt == SKIT_FMT_F32) wgrad_simt_kernel<1, 0><<<grid, 256, 0, st>>>(p);
        else wgrad_simt_kernel<1, 1><<<grid, 256, 0, st>>>(p);
        int rc = check_launch("wgrad_simt_kernel");
        if (rc) return rc;
        }
    }
    {
        int rc = unpack_wgrad(dwf, co_real, ci_real, k, dw, 1, layout, st, co, ci);
        if (rc) return rc;
    }
    if (dbias) return launch_dbias(dy, dy_org, ho, wo, co_real, dbias, st);
    return SKIT_OK;
}


extern "C" int skit_conv2d_wgrad_folded(const skit_operand* xf, int org, const skit_operand* dy, int dy_org, int k, int kw, int cp,
                                        int ho, int wo, float* scratch, float* dw, int co_real, int ci_real, void* stream) {
    SKIT_REQUIRE(xf && dy && scratch && dw && xf->p0 && xf->p1 && dy->p0 && dy->p1, "conv2d_wgrad_folded: null pointer");
    SKIT_REQUIRE(xf->fmt == SKIT_FMT_BF16X2 && dy->fmt == SKIT_FMT_BF16X2 && xf->c == 64 && dy->c % 64 == 0, "conv2d_wgrad_folded: needs bf16x2 operands, 64 folded channels");
    SKIT_REQUIRE(kw * cp <= 64 && ci_real <= cp && co_real <= dy->c && xf->n == dy->n, "conv2d_wgrad_folded: bad fold geometry");
    cudaStream_t st = as_stream(stream);
    int layout = 0;
    int rc = wgrad_tc(xf, org, dy, dy_org, k, 1, ho, wo, scratch, &layout, st, 1);
    if (rc) return rc;
    long long total = (long long)co_real * ci_real * k * kw;
    int blocks = (int)min((long long)148 * 8, cdivll(total, 256));
    unpack_wgrad_folded_kernel<<<blocks, 256, 0, st>>>(scratch, co_real, ci_real, k, kw, cp, dw, layout, dy->c);
    return check_launch("unpack_wgrad_folded_kernel");
}


/* Weight gradient of a thin-OUTPUT k x k layer (generator head 64 -> 5, k7) against the x-folded gradient operand that its input
 * gradient already uses:  dw[o][c][ky][kx] += sum_{y,X} dyf[y + k-1][X][(k-1-kx)*cp + o] * x[y + ky][X][c],  X over W + k - 1
 * columns — 7 row taps with the column taps inside the folded channels, instead of 49 passes over x. */
extern "C" int skit_conv2d_wgrad_dyfolded(const skit_operand* x, const skit_operand* dyf, int k, int cp, int ho, int wo,
                                          float* scratch, float* dw, int co_real, int ci_real, void* stream) {
    SKIT_REQUIRE(x && dyf && scratch && dw && x->p0 && x->p1 && dyf->p0 && dyf->p1, "conv2d_wgrad_dyfolded: null pointer");
    SKIT_REQUIRE(x->fmt == SKIT_FMT_BF16X2 && dyf->fmt == SKIT_FMT_BF16X2 && dyf->c == 64 && x->c % 64 == 0, "conv2d_wgrad_dyfolded: needs bf16x2 operands, 64 folded channels");
    SKIT_REQUIRE(k * cp <= 64 && co_real <= cp && ci_real <= x->c && x->n == dyf->n, "conv2d_wgrad_dyfolded: bad fold geometry");
    SKIT_REQUIRE(x->hp >= ho + k - 1 && x->wp >= wo + k - 1 && dyf->hp >= ho + k - 1 && dyf->wp >= wo + k - 1, "conv2d_wgrad_dyfolded: operands too small");
    cudaStream_t st = as_stream(stream);
    int layout = 0;
    // pixel domain: ho rows x (wo + k - 1) columns; x is read at (y + ky, X), the folded gradient at (y + k - 1, X)
    int rc = wgrad_tc(x, 0, dyf, 0, k, 1, ho, wo + k - 1, scratch, &layout, st, 1, k - 1);
    if (rc) return rc;
    long long total = (long long)co_real * ci_real * k * k;
    int blocks = (int)min((long long)148 * 8, cdivll(total, 256));
    unpack_wgrad_dyfolded_kernel<<<blocks, 256, 0, st>>>(scratch, co_real, ci_real, k, k, cp, dw, layout, x->c);
    return check_launch("unpack_wgrad_dyfolded_kernel");
}

extern "C" int skit_dbias_n(const skit_operand* dy, int dy_org, int ho, int wo, int nch, float* dbias, void* stream) {
    SKIT_REQUIRE(dy && dy->p0 && dbias && ho > 0 && wo > 0 && nch > 0 && nch <= dy->c, "dbias_n: bad arguments");
    SKIT_REQUIRE(dy_org + ho <= dy->hp && dy_org + wo <= dy->wp, "dbias_n: window exceeds the operand");
    return launch_dbias(dy, dy_org, ho, wo, nch, dbias, as_stream(stream));
}
