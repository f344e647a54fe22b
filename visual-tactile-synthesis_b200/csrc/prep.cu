// NHWC feature-map kernels between convolutions: statistics finalise, normalise + activation +
// residual + halo + bf16 split (forward), the two-phase norm/activation backward, and the
// antialiased blur resamplers with their adjoints.  All are HBM/L2-bandwidth bound: float4
// accesses along the channel axis, grid-stride loops sized in multiples of the SM count.
#include <cstdlib>
#include "skit_common.cuh"

namespace skit {

bool pdl_enabled() {
    static int on = -1;
    if (on < 0) { const char* e = getenv("SKIT_PDL"); on = e ? (atoi(e) != 0) : 0; }
    return on != 0;
}

constexpr int kSMs = 148;

static inline int grid_for(long long work, int threads, int max_per_sm = 8) {
    long long b = cdivll(work, threads);
    long long cap = (long long)kSMs * max_per_sm;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

// ------------------------------------------------------------------------------ stats finalise
__global__ void stats_finalize_kernel(const double* stats, int groups, int c, double count, float eps,
                                      float* mean_rstd, float* rmean, float* rvar, float momentum) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= groups * c) return;
    double mean = stats[2 * i] / count;
    double var = stats[2 * i + 1] / count - mean * mean;
    if (var < 0) var = 0;
    mean_rstd[2 * i] = (float)mean;
    mean_rstd[2 * i + 1] = (float)(1.0 / sqrt(var + (double)eps));
    if (rmean && i < c) {
        double unb = count > 1 ? var * count / (count - 1) : var;
        rmean[i] = (1.f - momentum) * rmean[i] + momentum * (float)mean;
        rvar[i] = (1.f - momentum) * rvar[i] + momentum * (float)unb;
    }
}

// ------------------------------------------------------------------------------ forward prep
struct PrepP {
    const float* raw; int n, h, w, c;
    const float* mr; int per_n;      // mean/rstd, indexed by n (instance) or 0 (batch); NULL = no norm
    const float* gamma; const float* beta;
    int act;
    const float* residual;
    float* out;                      // dense [n][h][w][c] or NULL
    float* o0; __nv_bfloat16* oh; __nv_bfloat16* ol; int fmt;  // operand planes or NULL
    int pad, pad_mode;
    int oc, ooff;                    // operand channel count and the channel offset this call fills (concat slices)
    // fused statistics finalise (rows kernel): the fp64 (sum, sum of squares) the producing conv's epilogue accumulated; every
    // thread derives mean / rstd of its own 4 channels in its prologue and one block per group writes them out for the backward
    const double* stats; double count; float eps; float* mr_out;
};

template <int VEC>
__global__ void __launch_bounds__(256) norm_act_pad_kernel(PrepP p) {
    const int hp = p.h + 2 * p.pad, wp = p.w + 2 * p.pad;
    const int cv = p.c / VEC;
    const long long total = (long long)p.n * hp * wp * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % cv) * VEC;
        long long t = i / cv;
        const int px = (int)(t % wp); t /= wp;
        const int py = (int)(t % hp);
        const int n = (int)(t / hp);
        const int sy = pad_src(py, p.pad, p.h, p.pad_mode), sx = pad_src(px, p.pad, p.w, p.pad_mode);
        float v[VEC];
        if (sy < 0 || sx < 0) {
#pragma unroll
            for (int j = 0; j < VEC; j++) v[j] = 0.f;
        } else {
            const long long src = (((long long)n * p.h + sy) * p.w + sx) * p.c + ch;
            if (VEC == 4) {
                float4 r = *reinterpret_cast<const float4*>(p.raw + src);
                v[0] = r.x; v[1 % VEC] = r.y; v[2 % VEC] = r.z; v[3 % VEC] = r.w;
            } else {
                v[0] = p.raw[src];
            }
#pragma unroll
            for (int j = 0; j < VEC; j++) {
                float x = v[j];
                if (p.mr) {
                    const float* mr = p.mr + ((long long)(p.per_n ? n : 0) * p.c + ch + j) * 2;
                    x = (x - mr[0]) * mr[1];
                }
                if (p.gamma) x = x * p.gamma[ch + j] + p.beta[ch + j];
                x = act_fwd(x, p.act);
                if (p.residual) x += p.residual[src + j];
                v[j] = x;
            }
            if (p.out && py - p.pad == sy && px - p.pad == sx) {
                if (VEC == 4) *reinterpret_cast<float4*>(p.out + src) = make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]);
                else p.out[src] = v[0];
            }
        }
        const long long dst = (((long long)n * hp + py) * wp + px) * p.oc + p.ooff + ch;
        if (p.o0 && p.fmt == SKIT_FMT_F32) {
            if (VEC == 4) *reinterpret_cast<float4*>(p.o0 + dst) = make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]);
            else p.o0[dst] = v[0];
        } else if (p.oh) {
            __nv_bfloat16 hi[VEC], lo[VEC];
#pragma unroll
            for (int j = 0; j < VEC; j++) split_bf16(v[j], hi[j], lo[j]);
            if (VEC == 4) {
                *reinterpret_cast<uint2*>(p.oh + dst) = *reinterpret_cast<uint2*>(hi);
                *reinterpret_cast<uint2*>(p.ol + dst) = *reinterpret_cast<uint2*>(lo);
            } else {
                p.oh[dst] = hi[0]; p.ol[dst] = lo[0];
            }
        }
    }
}


// Row-tiled variant for channel counts whose float4 lanes divide the 256-thread block (c = 4..1024, power-of-two lanes):
// a block owns whole padded rows of one image, a thread owns 4 fixed channels (its normalisation constants live in
// registers) and walks the row in steps of the block's pixel lanes — no integer divisions, no fp64, 4 independent
// 16-byte loads in flight per thread.
template <int FMT>
__global__ void __launch_bounds__(256) norm_act_pad_rows_kernel(PrepP p, int cv, int rows_per_block, int xseg) {
    pdl_trigger();
    pdl_wait();      // launched through launch_pdl: nothing above touches global memory
    const int hp = p.h + 2 * p.pad, wp = p.w + 2 * p.pad;
    const int xbeg = blockIdx.z * xseg, xend = min(wp, xbeg + xseg);   // blockIdx.z: segment of the row (small maps)
    const int n = blockIdx.y;
    const int cl = threadIdx.x % cv, pl = threadIdx.x / cv, PL = 256 / cv;
    const int ch = cl * 4;
    float mean[4] = {0, 0, 0, 0}, rstd[4] = {1, 1, 1, 1}, gam[4] = {1, 1, 1, 1}, bet[4] = {0, 0, 0, 0};
    const bool normed = p.mr != nullptr || p.stats != nullptr;
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (p.stats) {      // same arithmetic as stats_finalize_kernel
            const long long gi = (long long)(p.per_n ? n : 0) * p.c + ch + j;
            const double m = p.stats[2 * gi] / p.count;
            double var = p.stats[2 * gi + 1] / p.count - m * m;
            if (var < 0) var = 0;
            mean[j] = (float)m; rstd[j] = (float)(1.0 / sqrt(var + (double)p.eps));
            if (p.mr_out && blockIdx.x == 0 && blockIdx.z == 0 && pl == 0 && (p.per_n || n == 0)) {
                p.mr_out[2 * gi] = mean[j]; p.mr_out[2 * gi + 1] = rstd[j];
            }
        } else if (p.mr) {
            const float* mr = p.mr + ((long long)(p.per_n ? n : 0) * p.c + ch + j) * 2;
            mean[j] = mr[0]; rstd[j] = mr[1];
        }
        if (p.gamma) { gam[j] = p.gamma[ch + j]; bet[j] = p.beta[ch + j]; }
    }
    const int py_end = min(hp, (int)(blockIdx.x + 1) * rows_per_block);
    for (int py = blockIdx.x * rows_per_block; py < py_end; py++) {
        const int sy = pad_src(py, p.pad, p.h, p.pad_mode);
        const long long srow = ((long long)n * p.h + (sy < 0 ? 0 : sy)) * p.w;
        const long long drow = ((long long)n * hp + py) * wp;
        for (int px0 = xbeg + pl; px0 < xend; px0 += 4 * PL) {
            float4 r[4], res[4];
            int sx[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int px = px0 + u * PL;
                sx[u] = px < xend ? pad_src(px, p.pad, p.w, p.pad_mode) : -1;
                const bool ok = sy >= 0 && sx[u] >= 0;
                r[u] = ok ? *reinterpret_cast<const float4*>(p.raw + (srow + sx[u]) * p.c + ch) : make_float4(0, 0, 0, 0);
                if (p.residual) res[u] = ok ? *reinterpret_cast<const float4*>(p.residual + (srow + sx[u]) * p.c + ch) : make_float4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int px = px0 + u * PL;
                if (px >= xend) continue;
                const bool ok = sy >= 0 && sx[u] >= 0;
                float v[4] = {r[u].x, r[u].y, r[u].z, r[u].w};
                if (ok) {
                    const float rs[4] = {res[u].x, res[u].y, res[u].z, res[u].w};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        float x = v[j];
                        if (normed) x = (x - mean[j]) * rstd[j];
                        if (p.gamma) x = x * gam[j] + bet[j];
                        x = act_fwd(x, p.act);
                        if (p.residual) x += rs[j];
                        v[j] = x;
                    }
                    if (p.out && py - p.pad == sy && px - p.pad == sx[u])
                        *reinterpret_cast<float4*>(p.out + (srow + sx[u]) * p.c + ch) = make_float4(v[0], v[1], v[2], v[3]);
                }
                const long long dst = (drow + px) * p.oc + p.ooff + ch;
                if (FMT == SKIT_FMT_F32) {
                    if (p.o0) *reinterpret_cast<float4*>(p.o0 + dst) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
                    __nv_bfloat16 hi[4], lo[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) split_bf16(v[j], hi[j], lo[j]);
                    *reinterpret_cast<uint2*>(p.oh + dst) = *reinterpret_cast<uint2*>(hi);
                    *reinterpret_cast<uint2*>(p.ol + dst) = *reinterpret_cast<uint2*>(lo);
                }
            }
        }
    }
}

// ------------------------------------------------------------------------------ x-fold of a thin operand
// folded[n][y][x][kx*cp + c] = thin[n][y][x + kx][c]; one thread per (pixel, 4 folded channels).
struct FoldP {
    const float* t0; const __nv_bfloat16* th; const __nv_bfloat16* tl; int tfmt;
    int n, hp, wp, cp, kw, wf;
    __nv_bfloat16* oh; __nv_bfloat16* ol;
};
// grid: one block per (image, row) — no per-element 64-bit divisions; a thread produces 8 folded channels of one pixel (one 16-byte
// store per plane), the row's thin source (wp * cp floats, a few KB) stays in L1 across the kw shifted reads.
__global__ void __launch_bounds__(256) fold_x_kernel(FoldP p) {
    const int row = blockIdx.x;                       // b * hp + y
    const long long src_row = (long long)row * p.wp * p.cp;
    const long long dst_row = (long long)row * p.wf * 64;
    for (int i = threadIdx.x; i < p.wf * 8; i += 256) {
        const int x = i >> 3, j0 = (i & 7) * 8;
        __align__(16) __nv_bfloat16 hi[8], lo[8];
        int kx = j0 / p.cp, c = j0 - kx * p.cp;
#pragma unroll
        for (int q = 0; q < 8; q++) {
            float v = 0.f;
            if (kx < p.kw) {
                const long long a = src_row + (long long)(x + kx) * p.cp + c;
                v = p.tfmt == SKIT_FMT_F32 ? __ldg(p.t0 + a) : (__bfloat162float(p.th[a]) + __bfloat162float(p.tl[a]));
            }
            split_bf16(v, hi[q], lo[q]);
            if (++c == p.cp) { c = 0; kx++; }
        }
        const long long dst = dst_row + (long long)x * 64 + j0;
        *reinterpret_cast<uint4*>(p.oh + dst) = *reinterpret_cast<uint4*>(hi);
        *reinterpret_cast<uint4*>(p.ol + dst) = *reinterpret_cast<uint4*>(lo);
    }
}

// ------------------------------------------------------------------------------ backward phase A
// Layout of a `sums` buffer (doubles): [R][G][C][2] partial sums | [G][C] float2 finals (sum/count, packed in one double
// slot each) | 1 slot holding the int ticket counter.  The last CTA of the phase-A kernel adds the replicas up and writes
// the finals, so phase B reads two floats per channel and does no fp64 arithmetic at all.
__host__ __device__ inline long long sums_finals_offset(int groups, int c) { return (long long)SKIT_SUM_REPLICAS * groups * c * 2; }
__host__ __device__ inline long long sums_counter_offset(int groups, int c) { return sums_finals_offset(groups, c) + (long long)groups * c; }

__device__ inline void sums_last_block_finalize(double* sums, int groups, int c, double inv_count, unsigned total_blocks) {
    __shared__ unsigned s_ticket;
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) s_ticket = atomicAdd(reinterpret_cast<unsigned*>(sums + sums_counter_offset(groups, c)), 1u);
    __syncthreads();
    if (s_ticket != total_blocks - 1) return;
    __threadfence();
    float2* fin = reinterpret_cast<float2*>(sums + sums_finals_offset(groups, c));
    const long long rstride = (long long)groups * c * 2;
    for (int i = threadIdx.x; i < groups * c; i += blockDim.x) {
        double a0 = 0.0, a1 = 0.0;
        for (int r = 0; r < SKIT_SUM_REPLICAS; r++) {
            a0 += __ldcg(sums + r * rstride + 2 * i);
            a1 += __ldcg(sums + r * rstride + 2 * i + 1);
        }
        fin[i] = make_float2((float)(a0 * inv_count), (float)(a1 * inv_count));
    }
}

struct BwdAP {
    const float* dpad; int pad, pad_mode;
    const float* dadd;
    const float* dadd2;        // optional second dense gradient with the same slice geometry
    int dc0, dctot, dmask;     // dadd/dadd2 are channel slices [dc0, dc0+c) of [n][h][w][dctot]; dmask: multiply them by [pre > 0]
    const float* raw; int n, h, w, c;
    const float* mr; int per_n;
    const float* gamma; const float* beta;
    int act;
    float* g;
    double* sums;
    int chunk;  // pixels per block
    double inv_count;   // 1 / (elements per statistics group)
    float* dsum;        // optional second output: the incoming gradient itself, fold(dpad) + dadd (+ dadd2), before act'
};

// number of halo positions that alias source coordinate y on an axis of length len (reflect): returns
// their padded coordinates in out[], count as return value (always includes the interior one).
__device__ inline int fold_coords(int y, int pad, int len, int mode, int out[3]) {
    int cnt = 0;
    out[cnt++] = y + pad;
    if (mode == SKIT_PAD_REFLECT) {
        if (y >= 1 && y <= pad) out[cnt++] = pad - y;
        if (y <= len - 2 && y >= len - 1 - pad) out[cnt++] = pad + 2 * (len - 1) - y;
    }
    return cnt;
}

template <int VEC>
__global__ void __launch_bounds__(256) act_norm_bwd_reduce_kernel(BwdAP p) {
    __shared__ float red[256 * 2 * VEC];
    const int n = blockIdx.y;
    const int cv = p.c / VEC;
    const int lanes_c = cv < 256 ? cv : 256;
    const int PL = 256 / lanes_c;
    const int cl = threadIdx.x % lanes_c, pl = threadIdx.x / lanes_c;
    const int P = p.h * p.w;
    const int pbeg = blockIdx.x * p.chunk, pend = min(P, pbeg + p.chunk);
    const int hp = p.h + 2 * p.pad, wp = p.w + 2 * p.pad;
    for (int t = threadIdx.x; t < 256 * 2 * VEC; t += 256) red[t] = 0.f;
    __syncthreads();
    if (pl < PL) {
        for (int cvv = cl; cvv < cv; cvv += lanes_c) {
            const int ch = cvv * VEC;
            float mean[VEC], rstd[VEC], gam[VEC], bet[VEC];
#pragma unroll
            for (int j = 0; j < VEC; j++) {
                mean[j] = 0.f; rstd[j] = 1.f; gam[j] = 1.f; bet[j] = 0.f;
                if (p.mr) {
                    const float* mr = p.mr + ((long long)(p.per_n ? n : 0) * p.c + ch + j) * 2;
                    mean[j] = mr[0]; rstd[j] = mr[1];
                }
                if (p.gamma) { gam[j] = p.gamma[ch + j]; bet[j] = p.beta[ch + j]; }
            }
            float s0[VEC], s1[VEC];
#pragma unroll
            for (int j = 0; j < VEC; j++) { s0[j] = 0.f; s1[j] = 0.f; }
            for (int pix = pbeg + pl; pix < pend; pix += PL) {
                const int y = pix / p.w, x = pix - y * p.w;
                const long long src = (((long long)n * p.h + y) * p.w + x) * p.c + ch;
                float d[VEC];
                const long long dsrc = (((long long)n * p.h + y) * p.w + x) * p.dctot + p.dc0 + ch;
#pragma unroll
                for (int j = 0; j < VEC; j++) d[j] = (p.dadd ? p.dadd[dsrc + j] : 0.f) + (p.dadd2 ? p.dadd2[dsrc + j] : 0.f);
                if (p.dmask) {
#pragma unroll
                    for (int j = 0; j < VEC; j++) {
                        const float xh = ((p.raw ? p.raw[src + j] : 0.f) - mean[j]) * rstd[j];
                        if (!(xh * gam[j] + bet[j] > 0.f)) d[j] = 0.f;
                    }
                }
                if (p.dpad) {
                    int ys[3], xs[3];
                    const int ny = fold_coords(y, p.pad, p.h, p.pad_mode, ys);
                    const int nx = fold_coords(x, p.pad, p.w, p.pad_mode, xs);
                    for (int a = 0; a < ny; a++)
                        for (int b = 0; b < nx; b++) {
                            const long long q = (((long long)n * hp + ys[a]) * wp + xs[b]) * p.c + ch;
#pragma unroll
                            for (int j = 0; j < VEC; j++) d[j] += p.dpad[q + j];
                        }
                }
                if (p.dsum) {
#pragma unroll
                    for (int j = 0; j < VEC; j++) p.dsum[src + j] = d[j];
                }
                float gout[VEC];
#pragma unroll
                for (int j = 0; j < VEC; j++) {
                    const float xhat = ((p.raw ? p.raw[src + j] : 0.f) - mean[j]) * rstd[j];
                    const float pre = xhat * gam[j] + bet[j];
                    const float gg = act_grad(pre, p.act) * d[j];
                    gout[j] = gg;
                    s0[j] += gg; s1[j] += gg * xhat;
                }
                if (VEC == 4) *reinterpret_cast<float4*>(p.g + src) = make_float4(gout[0], gout[1 % VEC], gout[2 % VEC], gout[3 % VEC]);
                else p.g[src] = gout[0];
            }
            if (p.sums) {
#pragma unroll
                for (int j = 0; j < VEC; j++) {
                    atomicAdd(&red[(cl * VEC + j) * 2], s0[j]);
                    atomicAdd(&red[(cl * VEC + j) * 2 + 1], s1[j]);
                }
                // cv > lanes_c never happens for the supported channel counts (c <= 1024)
            }
        }
    }
    __syncthreads();
    if (p.sums) {
        for (int t = threadIdx.x; t < lanes_c * VEC; t += 256) {
            const long long rep = (long long)(blockIdx.x % SKIT_SUM_REPLICAS) * (p.per_n ? p.n : 1) * p.c * 2;
            double* dst = p.sums + rep + ((long long)(p.per_n ? n : 0) * p.c + t) * 2;
            atomicAdd(dst, (double)red[t * 2]);
            atomicAdd(dst + 1, (double)red[t * 2 + 1]);
        }
        sums_last_block_finalize(p.sums, p.per_n ? p.n : 1, p.c, p.inv_count, gridDim.x * gridDim.y);
    }
}


// Row-tiled variant of phase A (see norm_act_pad_rows_kernel): float4 loads of every stream, the halo fold without
// per-pixel divisions, per-thread register sums reduced across the block's pixel lanes, one fp64 atomic per channel per CTA.
template <int U, int MINB>
__global__ void __launch_bounds__(256, MINB) act_norm_bwd_reduce_rows_kernel(BwdAP p, int cv, int rows_per_block, int xseg) {
    pdl_trigger();
    pdl_wait();      // launched through launch_pdl: nothing above touches global memory
    __shared__ float red[256 * 8];
    const int n = blockIdx.y;
    const int cl = threadIdx.x % cv, pl = threadIdx.x / cv, PL = 256 / cv;
    const int ch = cl * 4;
    const int hp = p.h + 2 * p.pad, wp = p.w + 2 * p.pad;
    float mean[4] = {0, 0, 0, 0}, rstd[4] = {1, 1, 1, 1}, gam[4] = {1, 1, 1, 1}, bet[4] = {0, 0, 0, 0};
#pragma unroll
    for (int j = 0; j < 4; j++) {
        if (p.mr) {
            const float* mr = p.mr + ((long long)(p.per_n ? n : 0) * p.c + ch + j) * 2;
            mean[j] = mr[0]; rstd[j] = mr[1];
        }
        if (p.gamma) { gam[j] = p.gamma[ch + j]; bet[j] = p.beta[ch + j]; }
    }
    float s0[4] = {0, 0, 0, 0}, s1[4] = {0, 0, 0, 0};
    const int y_end = min(p.h, (int)(blockIdx.x + 1) * rows_per_block);
    for (int y = blockIdx.x * rows_per_block; y < y_end; y++) {
        // padded rows that alias source row y: the interior one, and for reflect padding the mirrored halo rows (-1 = none).
        // Scalars, not an index list: a list indexed by a run-time count lives in local memory and costs every pixel a store.
        const bool refl = p.dpad && p.pad_mode == SKIT_PAD_REFLECT;
        const int yc = y + p.pad;
        const int ylo = (refl && y >= 1 && y <= p.pad) ? p.pad - y : -1;
        const int yhi = (refl && y <= p.h - 2 && y >= p.h - 1 - p.pad) ? p.pad + 2 * (p.h - 1) - y : -1;
        const bool row_border = ylo >= 0 || yhi >= 0;
        const long long srow = ((long long)n * p.h + y) * p.w;
        const int xend = min(p.w, (int)(blockIdx.z + 1) * xseg);
        for (int x0 = blockIdx.z * xseg + pl; x0 < xend; x0 += U * PL) {
            // U pixels per pass: every independent 16-byte load is issued before the first use
            float4 d[U], rw[U], q0[U];
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int x = x0 + u * PL;
                const bool ok = x < xend;
                d[u] = make_float4(0, 0, 0, 0); rw[u] = d[u]; q0[u] = d[u];
                if (!ok) continue;
                if (p.dadd) {
                    const long long dsrc = (srow + x) * p.dctot + p.dc0 + ch;
                    d[u] = *reinterpret_cast<const float4*>(p.dadd + dsrc);
                    if (p.dadd2) {
                        const float4 e = *reinterpret_cast<const float4*>(p.dadd2 + dsrc);
                        d[u].x += e.x; d[u].y += e.y; d[u].z += e.z; d[u].w += e.w;
                    }
                }
                if (p.raw) rw[u] = *reinterpret_cast<const float4*>(p.raw + (srow + x) * p.c + ch);
                if (p.dpad) q0[u] = *reinterpret_cast<const float4*>(p.dpad + (((long long)n * hp + yc) * wp + x + p.pad) * p.c + ch);
            }
#pragma unroll
            for (int u = 0; u < U; u++) {
                const int x = x0 + u * PL;
                if (x >= xend) continue;
                const long long src = (srow + x) * p.c + ch;
                const float r4[4] = {rw[u].x, rw[u].y, rw[u].z, rw[u].w};
                float xhat[4], pre[4];
#pragma unroll
                for (int j = 0; j < 4; j++) { xhat[j] = (r4[j] - mean[j]) * rstd[j]; pre[j] = xhat[j] * gam[j] + bet[j]; }
                float dv[4] = {d[u].x, d[u].y, d[u].z, d[u].w};
                if (p.dmask) {
#pragma unroll
                    for (int j = 0; j < 4; j++) if (!(pre[j] > 0.f)) dv[j] = 0.f;
                }
                if (p.dpad) {
                    dv[0] += q0[u].x; dv[1] += q0[u].y; dv[2] += q0[u].z; dv[3] += q0[u].w;   // the interior position (yc, x + pad)
                    const int xlo = (refl && x >= 1 && x <= p.pad) ? p.pad - x : -1;
                    const int xhi = (refl && x <= p.w - 2 && x >= p.w - 1 - p.pad) ? p.pad + 2 * (p.w - 1) - x : -1;
                    if (row_border || xlo >= 0 || xhi >= 0) {     // border pixels also collect their reflected halo positions
                        const int ycand[3] = {yc, ylo, yhi}, xcand[3] = {x + p.pad, xlo, xhi};
#pragma unroll
                        for (int a = 0; a < 3; a++)
#pragma unroll
                            for (int b = 0; b < 3; b++) {
                                if ((a == 0 && b == 0) || ycand[a] < 0 || xcand[b] < 0) continue;
                                const float4 q = *reinterpret_cast<const float4*>(p.dpad + (((long long)n * hp + ycand[a]) * wp + xcand[b]) * p.c + ch);
                                dv[0] += q.x; dv[1] += q.y; dv[2] += q.z; dv[3] += q.w;
                            }
                    }
                }
                if (p.dsum) *reinterpret_cast<float4*>(p.dsum + src) = make_float4(dv[0], dv[1], dv[2], dv[3]);
                float go[4];
#pragma unroll
                for (int j = 0; j < 4; j++) {
                    go[j] = act_grad(pre[j], p.act) * dv[j];
                    s0[j] += go[j]; s1[j] = fmaf(go[j], xhat[j], s1[j]);
                }
                *reinterpret_cast<float4*>(p.g + src) = make_float4(go[0], go[1], go[2], go[3]);
            }
        }
    }
    if (p.sums) {
#pragma unroll
        for (int j = 0; j < 4; j++) { red[threadIdx.x * 8 + j] = s0[j]; red[threadIdx.x * 8 + 4 + j] = s1[j]; }
        __syncthreads();
        if (pl == 0) {
            float t0[4] = {0, 0, 0, 0}, t1[4] = {0, 0, 0, 0};
            for (int q = 0; q < PL; q++) {
#pragma unroll
                for (int j = 0; j < 4; j++) { t0[j] += red[(q * cv + cl) * 8 + j]; t1[j] += red[(q * cv + cl) * 8 + 4 + j]; }
            }
#pragma unroll
            for (int j = 0; j < 4; j++) {
                const long long rep = (long long)((blockIdx.x + 3 * blockIdx.z) % SKIT_SUM_REPLICAS) * (p.per_n ? p.n : 1) * p.c * 2;
                double* dst = p.sums + rep + ((long long)(p.per_n ? n : 0) * p.c + ch + j) * 2;
                atomicAdd(dst, (double)t0[j]);
                atomicAdd(dst + 1, (double)t1[j]);
            }
        }
        sums_last_block_finalize(p.sums, p.per_n ? p.n : 1, p.c, p.inv_count, gridDim.x * gridDim.y * gridDim.z);
    }
}

// ------------------------------------------------------------------------------ backward phase B
struct BwdBP {
    const float* g; const float* raw; int n, h, w, c;
    const float* mr; int per_n;
    const float* gamma;
    const double* sums; double inv_count;
    float* o0; __nv_bfloat16* oh; __nv_bfloat16* ol; int fmt;
    int pad;
    const float* extra;   // optional dense [n][h][w][c] gradient added to d_raw (a feature tap on the raw conv output)
};

template <int VEC>
__global__ void __launch_bounds__(256) norm_bwd_apply_kernel(BwdBP p) {
    const int hp = p.h + 2 * p.pad, wp = p.w + 2 * p.pad;
    const int cv = p.c / VEC;
    const long long total = (long long)p.n * hp * wp * cv;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % cv) * VEC;
        long long t = i / cv;
        const int px = (int)(t % wp); t /= wp;
        const int py = (int)(t % hp);
        const int n = (int)(t / hp);
        const int y = py - p.pad, x = px - p.pad;
        float v[VEC];
#pragma unroll
        for (int j = 0; j < VEC; j++) v[j] = 0.f;
        if (y >= 0 && y < p.h && x >= 0 && x < p.w) {
            const long long src = (((long long)n * p.h + y) * p.w + x) * p.c + ch;
#pragma unroll
            for (int j = 0; j < VEC; j++) {
                float gg = p.g[src + j];
                if (p.mr) {
                    const long long gi = ((long long)(p.per_n ? n : 0) * p.c + ch + j) * 2;
                    const float mean = p.mr[gi], rstd = p.mr[gi + 1];
                    const float xhat = (p.raw[src + j] - mean) * rstd;
                    const float2 fm = reinterpret_cast<const float2*>(p.sums + sums_finals_offset(p.per_n ? p.n : 1, p.c))[gi / 2];
                    const float m0 = fm.x, m1 = fm.y;
                    const float gam = p.gamma ? p.gamma[ch + j] : 1.f;
                    gg = gam * rstd * (gg - m0 - xhat * m1);
                }
                if (p.extra) gg += p.extra[src + j];
                v[j] = gg;
            }
        }
        const long long dst = (((long long)n * hp + py) * wp + px) * p.c + ch;
        if (p.fmt == SKIT_FMT_F32) {
            if (VEC == 4) *reinterpret_cast<float4*>(p.o0 + dst) = make_float4(v[0], v[1 % VEC], v[2 % VEC], v[3 % VEC]);
            else p.o0[dst] = v[0];
        } else {
            __nv_bfloat16 hi[VEC], lo[VEC];
#pragma unroll
            for (int j = 0; j < VEC; j++) split_bf16(v[j], hi[j], lo[j]);
            if (VEC == 4) {
                *reinterpret_cast<uint2*>(p.oh + dst) = *reinterpret_cast<uint2*>(hi);
                *reinterpret_cast<uint2*>(p.ol + dst) = *reinterpret_cast<uint2*>(lo);
            } else {
                p.oh[dst] = hi[0]; p.ol[dst] = lo[0];
            }
        }
    }
}


// Row-tiled variant (see norm_act_pad_rows_kernel): the per-channel terms of the norm backward (mean, rstd, gamma and
// the two reduced sums, converted from double ONCE per thread) stay in registers.
template <int FMT>
__global__ void __launch_bounds__(256) norm_bwd_apply_rows_kernel(BwdBP p, int cv, int rows_per_block, int xseg) {
    pdl_trigger();
    pdl_wait();      // launched through launch_pdl: nothing above touches global memory
    const int hp = p.h + 2 * p.pad, wp = p.w + 2 * p.pad;
    const int xbeg = blockIdx.z * xseg, xend = min(wp, xbeg + xseg);
    const int n = blockIdx.y;
    const int cl = threadIdx.x % cv, pl = threadIdx.x / cv, PL = 256 / cv;
    const int ch = cl * 4;
    float mean[4] = {0, 0, 0, 0}, rstd[4] = {1, 1, 1, 1}, gam[4] = {1, 1, 1, 1}, m0[4] = {0, 0, 0, 0}, m1[4] = {0, 0, 0, 0};
    if (p.mr) {
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const long long gi = ((long long)(p.per_n ? n : 0) * p.c + ch + j) * 2;
            mean[j] = p.mr[gi]; rstd[j] = p.mr[gi + 1];
            const float2 fm = reinterpret_cast<const float2*>(p.sums + sums_finals_offset(p.per_n ? p.n : 1, p.c))[gi / 2];
            m0[j] = fm.x; m1[j] = fm.y;
            if (p.gamma) gam[j] = p.gamma[ch + j];
        }
    }
    const int py_end = min(hp, (int)(blockIdx.x + 1) * rows_per_block);
    for (int py = blockIdx.x * rows_per_block; py < py_end; py++) {
        const int y = py - p.pad;
        const bool yin = y >= 0 && y < p.h;
        const long long srow = ((long long)n * p.h + (yin ? y : 0)) * p.w;
        const long long drow = ((long long)n * hp + py) * wp;
        for (int px0 = xbeg + pl; px0 < xend; px0 += 4 * PL) {
            float4 gg[4], rw[4];
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int x = px0 + u * PL - p.pad;
                const bool ok = yin && x >= 0 && x < p.w && px0 + u * PL < xend;
                gg[u] = ok ? *reinterpret_cast<const float4*>(p.g + (srow + x) * p.c + ch) : make_float4(0, 0, 0, 0);
                if (p.mr) rw[u] = ok ? *reinterpret_cast<const float4*>(p.raw + (srow + x) * p.c + ch) : make_float4(0, 0, 0, 0);
            }
#pragma unroll
            for (int u = 0; u < 4; u++) {
                const int px = px0 + u * PL;
                if (px >= xend) continue;
                const int x = px - p.pad;
                const bool ok = yin && x >= 0 && x < p.w;
                float v[4] = {gg[u].x, gg[u].y, gg[u].z, gg[u].w};
                if (ok && p.mr) {
                    const float r4[4] = {rw[u].x, rw[u].y, rw[u].z, rw[u].w};
#pragma unroll
                    for (int j = 0; j < 4; j++) {
                        const float xhat = (r4[j] - mean[j]) * rstd[j];
                        v[j] = gam[j] * rstd[j] * (v[j] - m0[j] - xhat * m1[j]);
                    }
                }
                if (ok && p.extra) {
                    const float4 e = *reinterpret_cast<const float4*>(p.extra + (srow + x) * p.c + ch);
                    v[0] += e.x; v[1] += e.y; v[2] += e.z; v[3] += e.w;
                }
                const long long dst = (drow + px) * p.c + ch;
                if (FMT == SKIT_FMT_F32) {
                    *reinterpret_cast<float4*>(p.o0 + dst) = make_float4(v[0], v[1], v[2], v[3]);
                } else {
                    __nv_bfloat16 hi[4], lo[4];
#pragma unroll
                    for (int j = 0; j < 4; j++) split_bf16(v[j], hi[j], lo[j]);
                    *reinterpret_cast<uint2*>(p.oh + dst) = *reinterpret_cast<uint2*>(hi);
                    *reinterpret_cast<uint2*>(p.ol + dst) = *reinterpret_cast<uint2*>(lo);
                }
            }
        }
    }
}

__global__ void bn_param_grad_kernel(const double* sums, int c, float count, float* dgamma, float* dbeta) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= c) return;
    const float2 fm = reinterpret_cast<const float2*>(sums + sums_finals_offset(1, c))[i];   // (S0, S1) / count
    if (dbeta) atomicAdd(dbeta + i, fm.x * count);
    if (dgamma) atomicAdd(dgamma + i, fm.y * count);
}

// ------------------------------------------------------------------------------ blur resamplers
// 1-D reflect(1) [1,2,1]/4 stride 2 down; bilinear-like [1,3,3,1]/4 stride 2 up with replicate edge.
// The four resamplers share one launch shape: blockIdx.y = row of the tensor being WRITTEN, blockIdx.z = image, threads stride over
// (x, 4 channels) of that row — no 64-bit div / mod per element.
__global__ void __launch_bounds__(256) blur_down_fwd_kernel(const float* __restrict__ x, int n, int h, int w, int c, float* __restrict__ y) {
    const int ho = h / 2, wo = w / 2, cv = c / 4;
    const int oy = blockIdx.y, b = blockIdx.z;
    const float f[3] = {0.25f, 0.5f, 0.25f};
    const float* rows[3];
#pragma unroll
    for (int a = 0; a < 3; a++) rows[a] = x + ((long long)b * h + pad_src(2 * oy + a, 1, h, SKIT_PAD_REFLECT)) * w * c;
    float* orow = y + ((long long)b * ho + oy) * wo * c;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < wo * cv; i += gridDim.x * blockDim.x) {
        const int ox = i / cv, ch = (i - ox * cv) * 4;
        float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int d = 0; d < 3; d++) {
            const long long off = (long long)pad_src(2 * ox + d, 1, w, SKIT_PAD_REFLECT) * c + ch;
#pragma unroll
            for (int a = 0; a < 3; a++) {
                const float wgt = f[a] * f[d];
                const float4 v = *reinterpret_cast<const float4*>(rows[a] + off);
                acc.x += wgt * v.x; acc.y += wgt * v.y; acc.z += wgt * v.z; acc.w += wgt * v.w;
            }
        }
        *reinterpret_cast<float4*>(orow + (long long)ox * c + ch) = acc;
    }
}

// Adjoint of blur_down in closed form (no tap lists in local memory): along one axis an even source index s feeds output s/2 with
// 0.5; an odd one feeds (s-1)/2 and (s+1)/2 with 0.25 each — the latter only while it exists, and s = 1 also collects the reflected
// padding sample (output 0, 0.25).  One block row per (image, source row): no 64-bit divisions per element.
__device__ __forceinline__ void blur_down_adj2(int s, int lo, int& i0, float& w0, int& i1, float& w1) {
    if ((s & 1) == 0) { i0 = s >> 1; w0 = 0.5f; i1 = i0; w1 = 0.f; return; }
    i0 = (s - 1) >> 1; w0 = s == 1 ? 0.5f : 0.25f;
    i1 = (s + 1) >> 1; w1 = i1 < lo ? 0.25f : 0.f;
    if (i1 >= lo) i1 = i0;
}

__global__ void __launch_bounds__(256) blur_down_bwd_kernel(const float* __restrict__ dy, int n, int h, int w, int c, float* __restrict__ dx) {
    const int ho = h / 2, wo = w / 2, cv = c / 4;
    const int y = blockIdx.y, b = blockIdx.z;
    int oy0, oy1; float wy0, wy1;
    blur_down_adj2(y, ho, oy0, wy0, oy1, wy1);
    const float* r0 = dy + ((long long)b * ho + oy0) * wo * c;
    const float* r1 = dy + ((long long)b * ho + oy1) * wo * c;
    float* orow = dx + ((long long)b * h + y) * w * c;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w * cv; i += gridDim.x * blockDim.x) {
        const int x = i / cv, ch = (i - x * cv) * 4;
        int ox0, ox1; float wx0, wx1;
        blur_down_adj2(x, wo, ox0, wx0, ox1, wx1);
        const float4 a00 = *reinterpret_cast<const float4*>(r0 + (long long)ox0 * c + ch);
        const float4 a01 = *reinterpret_cast<const float4*>(r0 + (long long)ox1 * c + ch);
        const float4 a10 = *reinterpret_cast<const float4*>(r1 + (long long)ox0 * c + ch);
        const float4 a11 = *reinterpret_cast<const float4*>(r1 + (long long)ox1 * c + ch);
        const float k00 = wy0 * wx0, k01 = wy0 * wx1, k10 = wy1 * wx0, k11 = wy1 * wx1;
        float4 acc;
        acc.x = k00 * a00.x + k01 * a01.x + k10 * a10.x + k11 * a11.x;
        acc.y = k00 * a00.y + k01 * a01.y + k10 * a10.y + k11 * a11.y;
        acc.z = k00 * a00.z + k01 * a01.z + k10 * a10.z + k11 * a11.z;
        acc.w = k00 * a00.w + k01 * a01.w + k10 * a10.w + k11 * a11.w;
        *reinterpret_cast<float4*>(orow + (long long)x * c + ch) = acc;
    }
}

__device__ inline void blur_up_taps(int u, int len, int src[2], float ws[2]) {
    const int m = u >> 1;
    if ((u & 1) == 0) { src[0] = max(m - 1, 0); ws[0] = 0.25f; src[1] = m; ws[1] = 0.75f; }
    else { src[0] = m; ws[0] = 0.75f; src[1] = min(m + 1, len - 1); ws[1] = 0.25f; }
}

__global__ void __launch_bounds__(256) blur_up_fwd_kernel(const float* __restrict__ x, int n, int h, int w, int c, float* __restrict__ y) {
    const int ho = 2 * h, wo = 2 * w, cv = c / 4;
    const int oy = blockIdx.y, b = blockIdx.z;
    int sy[2]; float wy[2];
    blur_up_taps(oy, h, sy, wy);
    const float* r0 = x + ((long long)b * h + sy[0]) * w * c;
    const float* r1 = x + ((long long)b * h + sy[1]) * w * c;
    float* orow = y + ((long long)b * ho + oy) * wo * c;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < wo * cv; i += gridDim.x * blockDim.x) {
        const int ox = i / cv, ch = (i - ox * cv) * 4;
        int sx[2]; float wx[2];
        blur_up_taps(ox, w, sx, wx);
        const float4 a00 = *reinterpret_cast<const float4*>(r0 + (long long)sx[0] * c + ch);
        const float4 a01 = *reinterpret_cast<const float4*>(r0 + (long long)sx[1] * c + ch);
        const float4 a10 = *reinterpret_cast<const float4*>(r1 + (long long)sx[0] * c + ch);
        const float4 a11 = *reinterpret_cast<const float4*>(r1 + (long long)sx[1] * c + ch);
        const float k00 = wy[0] * wx[0], k01 = wy[0] * wx[1], k10 = wy[1] * wx[0], k11 = wy[1] * wx[1];
        float4 acc;
        acc.x = k00 * a00.x + k01 * a01.x + k10 * a10.x + k11 * a11.x;
        acc.y = k00 * a00.y + k01 * a01.y + k10 * a10.y + k11 * a11.y;
        acc.z = k00 * a00.z + k01 * a01.z + k10 * a10.z + k11 * a11.z;
        acc.w = k00 * a00.w + k01 * a01.w + k10 * a10.w + k11 * a11.w;
        *reinterpret_cast<float4*>(orow + (long long)ox * c + ch) = acc;
    }
}

// Adjoint of blur_up in closed form: source m feeds outputs 2m-1 (0.25, if m > 0), 2m (0.75, +0.25 at the clamped low edge),
// 2m+1 (0.75, +0.25 at the clamped high edge) and 2m+2 (0.25, if m + 1 < len).
__device__ __forceinline__ void blur_up_adj4(int m, int len, int idx[4], float wt[4]) {
    idx[0] = m > 0 ? 2 * m - 1 : 0;            wt[0] = m > 0 ? 0.25f : 0.f;
    idx[1] = 2 * m;                            wt[1] = m == 0 ? 1.0f : 0.75f;
    idx[2] = 2 * m + 1;                        wt[2] = m == len - 1 ? 1.0f : 0.75f;
    idx[3] = m + 1 < len ? 2 * m + 2 : 2 * m;  wt[3] = m + 1 < len ? 0.25f : 0.f;
}

__global__ void __launch_bounds__(256) blur_up_bwd_kernel(const float* __restrict__ dy, int n, int h, int w, int c, float* __restrict__ dx) {
    const int ho = 2 * h, wo = 2 * w, cv = c / 4;
    const int y = blockIdx.y, b = blockIdx.z;
    int oys[4]; float wy[4];
    blur_up_adj4(y, h, oys, wy);
    const float* rows[4];
#pragma unroll
    for (int a = 0; a < 4; a++) rows[a] = dy + ((long long)b * ho + oys[a]) * wo * c;
    float* orow = dx + ((long long)b * h + y) * w * c;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < w * cv; i += gridDim.x * blockDim.x) {
        const int x = i / cv, ch = (i - x * cv) * 4;
        int oxs[4]; float wx[4];
        blur_up_adj4(x, w, oxs, wx);
        float4 acc = make_float4(0, 0, 0, 0);
#pragma unroll
        for (int a = 0; a < 4; a++)
#pragma unroll
            for (int d = 0; d < 4; d++) {
                const float wgt = wy[a] * wx[d];
                const float4 v = *reinterpret_cast<const float4*>(rows[a] + (long long)oxs[d] * c + ch);
                acc.x += wgt * v.x; acc.y += wgt * v.y; acc.z += wgt * v.z; acc.w += wgt * v.w;
            }
        *reinterpret_cast<float4*>(orow + (long long)x * c + ch) = acc;
    }
}

// channel counts the row-tiled kernels take: float4 lanes c/4 must divide the 256-thread block
static inline bool rows_layout_ok(int c) { return c % 4 == 0 && c / 4 <= 256 && 256 % (c / 4) == 0; }
// rows per block so that the grid has a few CTAs per SM
static inline int rows_per_block_for(int rows, int n) {
    const int want_blocks = max(1, (kSMs * 6) / max(1, n));
    return max(1, cdiv(rows, want_blocks));
}
// when one row per block still leaves the grid short (small maps), split each row into segments (grid.z)
static inline int xseg_for(int rows, int width, int n, int pixel_lanes) {
    const int want_blocks = max(1, (kSMs * 6) / max(1, n));
    const int rpb = rows_per_block_for(rows, n);
    const int row_blocks = cdiv(rows, rpb);
    int segs = max(1, want_blocks / row_blocks);
    int seg = max(4 * pixel_lanes, cdiv(width, segs));   // at least one full unrolled pass per thread
    return min(seg, width);
}

static int check_operand(const skit_operand* op, int n, int h, int w, int c, int pad, const char* who) {
    if (!op) return SKIT_OK;
    if (!op->p0 || (op->fmt == SKIT_FMT_BF16X2 && !op->p1) || (op->fmt != SKIT_FMT_F32 && op->fmt != SKIT_FMT_BF16X2)) {
        set_error("%s: bad operand pointers/format", who);
        return SKIT_ERR_INVALID;
    }
    if (op->n != n || op->hp != h + 2 * pad || op->wp != w + 2 * pad || op->c != c) {
        set_error("%s: operand dims [%d,%d,%d,%d] do not match [%d,%d+2*%d,%d+2*%d,%d]", who, op->n, op->hp, op->wp, op->c, n, h, pad, w, pad, c);
        return SKIT_ERR_INVALID;
    }
    return SKIT_OK;
}

}  // namespace skit

using namespace skit;

extern "C" int skit_stats_finalize(const double* stats, int groups, int c, double count, float eps,
                                   float* mean_rstd, float* running_mean, float* running_var, float momentum, void* stream) {
    SKIT_REQUIRE(stats && mean_rstd && groups > 0 && c > 0 && count > 0, "stats_finalize: bad arguments");
    SKIT_REQUIRE((running_mean == nullptr) == (running_var == nullptr), "stats_finalize: running stats must come in pairs");
    SKIT_REQUIRE(!running_mean || groups == 1, "stats_finalize: running stats only for batch statistics (groups == 1)");
    stats_finalize_kernel<<<cdiv(groups * c, 128), 128, 0, as_stream(stream)>>>(stats, groups, c, count, eps, mean_rstd, running_mean, running_var, momentum);
    return check_launch("stats_finalize_kernel");
}

extern "C" int skit_fold_x_operand(const skit_operand* thin, int kw, const skit_operand* folded, void* stream) {
    SKIT_REQUIRE(thin && folded && thin->p0 && folded->p0 && folded->p1 && folded->fmt == SKIT_FMT_BF16X2, "fold_x_operand: bad operands (bf16x2 destination required)");
    SKIT_REQUIRE(thin->fmt == SKIT_FMT_F32 || (thin->fmt == SKIT_FMT_BF16X2 && thin->p1), "fold_x_operand: bad source format");
    SKIT_REQUIRE(kw >= 1 && kw * thin->c <= 64 && folded->c == 64, "fold_x_operand: kw*c = %d*%d must fit the 64 folded channels", kw, thin->c);
    SKIT_REQUIRE(folded->n == thin->n && folded->hp == thin->hp && folded->wp == thin->wp - kw + 1, "fold_x_operand: destination dims mismatch");
    FoldP p{};
    p.t0 = (const float*)thin->p0; p.th = (const __nv_bfloat16*)thin->p0; p.tl = (const __nv_bfloat16*)thin->p1; p.tfmt = thin->fmt;
    p.n = thin->n; p.hp = thin->hp; p.wp = thin->wp; p.cp = thin->c; p.kw = kw; p.wf = folded->wp;
    p.oh = (__nv_bfloat16*)folded->p0; p.ol = (__nv_bfloat16*)folded->p1;
    fold_x_kernel<<<p.n * p.hp, 256, 0, as_stream(stream)>>>(p);
    return check_launch("fold_x_kernel");
}

static int norm_act_pad_impl(const float* raw, int n, int h, int w, int c,
                             const float* mean_rstd, const double* stats, double count, float eps, float* mr_out,
                             int norm_mode, const float* gamma, const float* beta,
                             int act, const float* residual, float* out,
                             const skit_operand* op, int c_off, int pad, int pad_mode, void* stream) {
    SKIT_REQUIRE(raw && n > 0 && h > 0 && w > 0 && c > 0, "norm_act_pad: bad arguments");
    SKIT_REQUIRE(out || op, "norm_act_pad: nothing to write");
    SKIT_REQUIRE((norm_mode == SKIT_NORM_NONE) == (mean_rstd == nullptr && stats == nullptr), "norm_act_pad: mean_rstd / stats must be given iff norm_mode != none");
    if (stats && !(rows_layout_ok(c) && (!op || (op->c % 4 == 0 && c_off % 4 == 0)))) {
        // layouts the row-tiled kernel does not cover: finalise with the stand-alone kernel, then the generic pass
        const int groups = norm_mode == SKIT_NORM_INSTANCE ? n : 1;
        stats_finalize_kernel<<<cdiv(groups * c, 128), 128, 0, as_stream(stream)>>>(stats, groups, c, count, eps, mr_out, nullptr, nullptr, 0.f);
        int rc0 = check_launch("stats_finalize_kernel");
        if (rc0) return rc0;
        mean_rstd = mr_out; stats = nullptr;
    }
    SKIT_REQUIRE((gamma == nullptr) == (beta == nullptr), "norm_act_pad: gamma/beta must come in pairs");
    SKIT_REQUIRE(pad >= 0 && (pad_mode != SKIT_PAD_REFLECT || (pad < h && pad < w)), "norm_act_pad: reflect pad %d too large for %dx%d", pad, h, w);
    if (!op) pad = 0;
    const int oc = op ? op->c : c;
    SKIT_REQUIRE(c_off >= 0 && c_off + c <= oc, "norm_act_pad: channel slice [%d, %d) exceeds the operand's %d channels", c_off, c_off + c, oc);
    int rc = check_operand(op, n, h, w, oc, pad, "norm_act_pad");
    if (rc) return rc;
    PrepP p{};
    p.raw = raw; p.n = n; p.h = h; p.w = w; p.c = c;
    p.mr = mean_rstd; p.per_n = norm_mode == SKIT_NORM_INSTANCE; p.gamma = gamma; p.beta = beta;
    p.act = act; p.residual = residual; p.out = out;
    if (op) {
        p.fmt = op->fmt;
        if (op->fmt == SKIT_FMT_F32) p.o0 = (float*)op->p0;
        else { p.oh = (__nv_bfloat16*)op->p0; p.ol = (__nv_bfloat16*)op->p1; }
    }
    p.pad = pad; p.pad_mode = pad_mode; p.oc = oc; p.ooff = c_off;
    p.stats = stats; p.count = count; p.eps = eps; p.mr_out = mr_out;
    const long long pix = (long long)n * (h + 2 * pad) * (w + 2 * pad);
    if (rows_layout_ok(c) && (!op || (oc % 4 == 0 && c_off % 4 == 0))) {
        const int hp = h + 2 * pad, wp = w + 2 * pad;
        const int rpb = rows_per_block_for(hp, n);
        const int xs = xseg_for(hp, wp, n, 256 / (c / 4));
        dim3 grid(cdiv(hp, rpb), n, cdiv(wp, xs));
        if (!op || op->fmt == SKIT_FMT_F32) launch_pdl(norm_act_pad_rows_kernel<SKIT_FMT_F32>, grid, 256, 0, as_stream(stream), p, c / 4, rpb, xs);
        else launch_pdl(norm_act_pad_rows_kernel<SKIT_FMT_BF16X2>, grid, 256, 0, as_stream(stream), p, c / 4, rpb, xs);
        return check_launch("norm_act_pad_rows_kernel");
    }
    if (c % 4 == 0 && oc % 4 == 0 && c_off % 4 == 0) norm_act_pad_kernel<4><<<grid_for(pix * (c / 4), 256), 256, 0, as_stream(stream)>>>(p);
    else norm_act_pad_kernel<1><<<grid_for(pix * c, 256), 256, 0, as_stream(stream)>>>(p);
    return check_launch("norm_act_pad_kernel");
}

extern "C" int skit_norm_act_pad_ex(const float* raw, int n, int h, int w, int c,
                                    const float* mean_rstd, int norm_mode, const float* gamma, const float* beta,
                                    int act, const float* residual, float* out,
                                    const skit_operand* op, int c_off, int pad, int pad_mode, void* stream) {
    return norm_act_pad_impl(raw, n, h, w, c, mean_rstd, nullptr, 0.0, 0.f, nullptr, norm_mode, gamma, beta, act, residual, out, op, c_off,
                             pad, pad_mode, stream);
}

extern "C" int skit_norm_act_pad_stats(const float* raw, int n, int h, int w, int c,
                                       const double* stats, double count, float eps, float* mean_rstd_out,
                                       int norm_mode, const float* gamma, const float* beta,
                                       int act, const float* residual, float* out,
                                       const skit_operand* op, int c_off, int pad, int pad_mode, void* stream) {
    SKIT_REQUIRE(stats && mean_rstd_out && count > 0 && norm_mode != SKIT_NORM_NONE, "norm_act_pad_stats: statistics, their count and the mean/rstd output are required");
    return norm_act_pad_impl(raw, n, h, w, c, nullptr, stats, count, eps, mean_rstd_out, norm_mode, gamma, beta, act, residual, out, op, c_off,
                             pad, pad_mode, stream);
}

extern "C" int skit_norm_act_pad(const float* raw, int n, int h, int w, int c,
                                 const float* mean_rstd, int norm_mode, const float* gamma, const float* beta,
                                 int act, const float* residual, float* out,
                                 const skit_operand* op, int pad, int pad_mode, void* stream) {
    SKIT_REQUIRE(!op || op->c == c, "norm_act_pad: operand has %d channels, expected %d", op ? op->c : 0, c);
    return skit_norm_act_pad_ex(raw, n, h, w, c, mean_rstd, norm_mode, gamma, beta, act, residual, out, op, 0, pad, pad_mode, stream);
}

extern "C" int skit_act_norm_bwd_reduce_ex2(const float* dpad, int pad, int pad_mode,
                                            const float* dadd, const float* dadd2, int dadd_c0, int dadd_ctot, int dadd_relu_mask,
                                            const float* raw, int n, int h, int w, int c,
                                            const float* mean_rstd, int norm_mode, const float* gamma, const float* beta,
                                            int act, float* g, double* sums, float* dsum, void* stream);

extern "C" int skit_act_norm_bwd_reduce_ex(const float* dpad, int pad, int pad_mode,
                                           const float* dadd, const float* dadd2, int dadd_c0, int dadd_ctot, int dadd_relu_mask,
                                           const float* raw, int n, int h, int w, int c,
                                           const float* mean_rstd, int norm_mode, const float* gamma, const float* beta,
                                           int act, float* g, double* sums, void* stream) {
    return skit_act_norm_bwd_reduce_ex2(dpad, pad, pad_mode, dadd, dadd2, dadd_c0, dadd_ctot, dadd_relu_mask, raw, n, h, w, c,
                                        mean_rstd, norm_mode, gamma, beta, act, g, sums, nullptr, stream);
}

extern "C" int skit_act_norm_bwd_reduce_ex2(const float* dpad, int pad, int pad_mode,
                                            const float* dadd, const float* dadd2, int dadd_c0, int dadd_ctot, int dadd_relu_mask,
                                            const float* raw, int n, int h, int w, int c,
                                            const float* mean_rstd, int norm_mode, const float* gamma, const float* beta,
                                            int act, float* g, double* sums, float* dsum, void* stream) {
    SKIT_REQUIRE(g && (dpad || dadd) && n > 0 && h > 0 && w > 0 && c > 0, "act_norm_bwd_reduce: bad arguments");
    SKIT_REQUIRE(raw || (norm_mode == SKIT_NORM_NONE && act == SKIT_ACT_NONE && !dadd_relu_mask), "act_norm_bwd_reduce: raw required unless norm, act and mask are all none");
    SKIT_REQUIRE((norm_mode == SKIT_NORM_NONE) == (mean_rstd == nullptr), "act_norm_bwd_reduce: mean_rstd must be given iff norm_mode != none");
    SKIT_REQUIRE(norm_mode == SKIT_NORM_NONE || sums, "act_norm_bwd_reduce: sums required with a norm");
    SKIT_REQUIRE(pad_mode == SKIT_PAD_ZERO || pad_mode == SKIT_PAD_REFLECT, "act_norm_bwd_reduce: unsupported pad mode");
    SKIT_REQUIRE((gamma == nullptr) == (beta == nullptr), "act_norm_bwd_reduce: gamma/beta must come in pairs");
    SKIT_REQUIRE(!dadd2 || dadd, "act_norm_bwd_reduce: dadd2 without dadd");
    SKIT_REQUIRE(dadd_c0 >= 0 && dadd_c0 + c <= dadd_ctot, "act_norm_bwd_reduce: dadd slice [%d, %d) exceeds %d channels", dadd_c0, dadd_c0 + c, dadd_ctot);
    const int vec = (c % 4 == 0) ? 4 : 1;
    SKIT_REQUIRE(c / vec <= 256, "act_norm_bwd_reduce: channel count %d too large", c);
    BwdAP p{};
    p.dpad = dpad; p.pad = dpad ? pad : 0; p.pad_mode = pad_mode; p.dadd = dadd; p.dadd2 = dadd2;
    p.dc0 = dadd_c0; p.dctot = dadd_ctot; p.dmask = dadd_relu_mask;
    p.raw = raw; p.n = n; p.h = h; p.w = w; p.c = c;
    p.mr = mean_rstd; p.per_n = norm_mode == SKIT_NORM_INSTANCE; p.gamma = gamma; p.beta = beta;
    p.act = act; p.g = g; p.sums = sums; p.dsum = dsum;
    p.inv_count = 1.0 / ((double)h * w * (norm_mode == SKIT_NORM_BATCH ? n : 1));
    const int P = h * w;
    if (rows_layout_ok(c) && dadd_c0 % 4 == 0 && dadd_ctot % 4 == 0) {
        // pixels per pass / resident blocks: 2 pixels at 3 blocks per SM measured 47.8 us at 192 x 192 x 256 against 56.5 us for
        // 4 pixels at 2 blocks per SM (profiles/r02f_bench_prep.txt); SKIT_REDUCE_UNROLL = 4 | 2 | 1 selects the variant.
        // grid: the kernel pays a per-block tail (shared-memory reduction, 8 fp64 atomics
        // per channel lane, fence + ticket): `per_sm` blocks per SM in total (SKIT_REDUCE_BLOCKS_PER_SM, default 6)
        static int per_sm = -1;
        if (per_sm < 0) { const char* e = getenv("SKIT_REDUCE_BLOCKS_PER_SM"); per_sm = e ? atoi(e) : 6; if (per_sm < 1) per_sm = 6; }
        const int want_blocks = max(1, (kSMs * per_sm) / max(1, n));
        const int rpb = max(1, cdiv(h, want_blocks));
        const int row_blocks = cdiv(h, rpb);
        const int PLr = 256 / (c / 4);
        int xs = max(4 * PLr, cdiv(w, max(1, want_blocks / row_blocks)));
        if (xs > w) xs = w;
        dim3 grid(row_blocks, n, cdiv(w, xs));
        static int unroll = -1;
        if (unroll < 0) { const char* e = getenv("SKIT_REDUCE_UNROLL"); unroll = e ? atoi(e) : 2; }
        if (unroll == 2) launch_pdl(act_norm_bwd_reduce_rows_kernel<2, 3>, grid, 256, 0, as_stream(stream), p, c / 4, rpb, xs);
        else if (unroll == 1) launch_pdl(act_norm_bwd_reduce_rows_kernel<1, 4>, grid, 256, 0, as_stream(stream), p, c / 4, rpb, xs);
        else launch_pdl(act_norm_bwd_reduce_rows_kernel<4, 2>, grid, 256, 0, as_stream(stream), p, c / 4, rpb, xs);
        return check_launch("act_norm_bwd_reduce_rows_kernel");
    }
    const int lanes_c = min(c / vec, 256), PL = 256 / lanes_c;
    int blocks_per_n = max(1, min(cdiv(P, PL * 4), cdiv(kSMs * 8, n)));
    p.chunk = cdiv(P, blocks_per_n);
    dim3 grid(cdiv(P, p.chunk), n);
    if (vec == 4) act_norm_bwd_reduce_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(p);
    else act_norm_bwd_reduce_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(p);
    return check_launch("act_norm_bwd_reduce_kernel");
}

extern "C" int skit_act_norm_bwd_reduce(const float* dpad, int pad, int pad_mode, const float* dadd,
                                        const float* raw, int n, int h, int w, int c,
                                        const float* mean_rstd, int norm_mode, const float* gamma, const float* beta,
                                        int act, float* g, double* sums, void* stream) {
    return skit_act_norm_bwd_reduce_ex(dpad, pad, pad_mode, dadd, nullptr, 0, c, 0, raw, n, h, w, c, mean_rstd, norm_mode,
                                       gamma, beta, act, g, sums, stream);
}

extern "C" int skit_norm_bwd_apply_ex(const float* g, const float* raw, int n, int h, int w, int c,
                                      const float* mean_rstd, int norm_mode, const float* gamma,
                                      const double* sums, double count, float* dgamma, float* dbeta,
                                      const float* extra, const skit_operand* op, int pad, void* stream);

extern "C" int skit_norm_bwd_apply(const float* g, const float* raw, int n, int h, int w, int c,
                                   const float* mean_rstd, int norm_mode, const float* gamma,
                                   const double* sums, double count, float* dgamma, float* dbeta,
                                   const skit_operand* op, int pad, void* stream) {
    return skit_norm_bwd_apply_ex(g, raw, n, h, w, c, mean_rstd, norm_mode, gamma, sums, count, dgamma, dbeta, nullptr, op, pad, stream);
}

extern "C" int skit_norm_bwd_apply_ex(const float* g, const float* raw, int n, int h, int w, int c,
                                      const float* mean_rstd, int norm_mode, const float* gamma,
                                      const double* sums, double count, float* dgamma, float* dbeta,
                                      const float* extra, const skit_operand* op, int pad, void* stream) {
    SKIT_REQUIRE(g && op && n > 0 && h > 0 && w > 0 && c > 0 && pad >= 0, "norm_bwd_apply: bad arguments");
    SKIT_REQUIRE((norm_mode == SKIT_NORM_NONE) == (mean_rstd == nullptr), "norm_bwd_apply: mean_rstd must be given iff norm_mode != none");
    SKIT_REQUIRE(norm_mode == SKIT_NORM_NONE || (sums && raw && count > 0), "norm_bwd_apply: sums/raw/count required with a norm");
    int rc = check_operand(op, n, h, w, c, pad, "norm_bwd_apply");
    if (rc) return rc;
    BwdBP p{};
    p.g = g; p.raw = raw; p.n = n; p.h = h; p.w = w; p.c = c;
    p.mr = mean_rstd; p.per_n = norm_mode == SKIT_NORM_INSTANCE; p.gamma = gamma;
    p.sums = sums; p.inv_count = count > 0 ? 1.0 / count : 0.0;
    p.extra = extra;
    p.fmt = op->fmt;
    if (op->fmt == SKIT_FMT_F32) p.o0 = (float*)op->p0;
    else { p.oh = (__nv_bfloat16*)op->p0; p.ol = (__nv_bfloat16*)op->p1; }
    p.pad = pad;
    const long long pix = (long long)n * (h + 2 * pad) * (w + 2 * pad);
    if (rows_layout_ok(c)) {
        const int hp = h + 2 * pad, wp = w + 2 * pad;
        const int rpb = rows_per_block_for(hp, n);
        const int xs = xseg_for(hp, wp, n, 256 / (c / 4));
        dim3 grid(cdiv(hp, rpb), n, cdiv(wp, xs));
        if (op->fmt == SKIT_FMT_F32) launch_pdl(norm_bwd_apply_rows_kernel<SKIT_FMT_F32>, grid, 256, 0, as_stream(stream), p, c / 4, rpb, xs);
        else launch_pdl(norm_bwd_apply_rows_kernel<SKIT_FMT_BF16X2>, grid, 256, 0, as_stream(stream), p, c / 4, rpb, xs);
    } else if (c % 4 == 0) norm_bwd_apply_kernel<4><<<grid_for(pix * (c / 4), 256), 256, 0, as_stream(stream)>>>(p);
    else norm_bwd_apply_kernel<1><<<grid_for(pix * c, 256), 256, 0, as_stream(stream)>>>(p);
    rc = check_launch("norm_bwd_apply_kernel");
    if (rc) return rc;
    if (norm_mode == SKIT_NORM_BATCH && (dgamma || dbeta)) {
        bn_param_grad_kernel<<<cdiv(c, 128), 128, 0, as_stream(stream)>>>(sums, c, (float)count, dgamma, dbeta);
        return check_launch("bn_param_grad_kernel");
    }
    return SKIT_OK;
}

// rows_out x width_out: the tensor the kernel writes (one block row per output row, a few CTAs per SM in total)
#define SKIT_BLUR_ENTRY(name, kernel, rows_out, width_out, cond)                                         \
    extern "C" int name(const float* a, int n, int h, int w, int c, float* b, void* stream) {           \
        SKIT_REQUIRE(a && b && n > 0 && h > 1 && w > 1 && c > 0 && c % 4 == 0 && (cond),                 \
                     #name ": bad arguments (n=%d h=%d w=%d c=%d)", n, h, w, c);                         \
        const int rows = (rows_out), per_row = (width_out) * (c / 4);                                    \
        SKIT_REQUIRE(rows <= 65535 && n <= 65535, #name ": map too tall / batch too large for the row grid"); \
        int bx = cdiv(per_row, 256);                                                                     \
        const int want = cdiv(kSMs * 8, rows * n);                                                       \
        if (bx > want) bx = want < 1 ? 1 : want;                                                         \
        kernel<<<dim3(bx, rows, n), 256, 0, as_stream(stream)>>>(a, n, h, w, c, b);                      \
        return check_launch(#kernel);                                                                    \
    }

SKIT_BLUR_ENTRY(skit_blur_down_fwd, blur_down_fwd_kernel, h / 2, w / 2, (h % 2 == 0 && w % 2 == 0))
SKIT_BLUR_ENTRY(skit_blur_down_bwd, blur_down_bwd_kernel, h, w, (h % 2 == 0 && w % 2 == 0))
SKIT_BLUR_ENTRY(skit_blur_up_fwd, blur_up_fwd_kernel, 2 * h, 2 * w, true)
SKIT_BLUR_ENTRY(skit_blur_up_bwd, blur_up_bwd_kernel, h, w, true)
