// Image-level (NCHW fp32 planar, 1-7 channels) kernels of the train step: layout staging into
// haloed NHWC operands, the generator head, DiffAugment, avg-pool pyramid, patch gather/scatter,
// GAN/L1 losses, Adam, and the PatchNCE sampling + loss.  All HBM-bound; coalesced along W.
#include "skit_common.cuh"

#include <cstdarg>

namespace skit {

static thread_local char g_err[512] = "";

void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int check_launch(const char* what) {
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) {
        set_error("%s: launch failed: %s", what, cudaGetErrorString(e));
        return SKIT_ERR_CUDA;
    }
    return SKIT_OK;
}

static inline int grid_for(long long work, int threads, int max_per_sm = 8) {
    long long b = cdivll(work, threads);
    long long cap = 148LL * max_per_sm;
    return (int)(b < 1 ? 1 : (b > cap ? cap : b));
}

struct Srcs4 {
    const float* p[4];
    int ch[4];
    int coff[4];
    int n;
};

// ------------------------------------------------------------------------------ layout staging
// cop >= ctot: channels beyond the sources are zero (channel padding for the tensor-core path, whose TMA rows must be
// 16-byte multiples).  fmt selects fp32 or bf16 hi/lo planes.
__global__ void __launch_bounds__(256) nchw_cat_to_operand_kernel(Srcs4 s, int n, int h, int w, int ctot, int cop, int fmt,
                                                                  float* dst, __nv_bfloat16* dh, __nv_bfloat16* dl, int pad, int pad_mode) {
    const int hp = h + 2 * pad, wp = w + 2 * pad;
    const long long total = (long long)n * hp * wp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int px = (int)(i % wp);
        long long t = i / wp;
        const int py = (int)(t % hp);
        const int b = (int)(t / hp);
        const int sy = pad_src(py, pad, h, pad_mode), sx = pad_src(px, pad, w, pad_mode);
        int c = 0;
        for (int k = 0; k < s.n; k++)
            for (int j = 0; j < s.ch[k]; j++, c++) {
                const float v = (sy < 0 || sx < 0) ? 0.f : __ldg(s.p[k] + (((long long)b * s.ch[k] + j) * h + sy) * w + sx);
                if (fmt == SKIT_FMT_F32) dst[i * cop + c] = v;
                else { __nv_bfloat16 hi, lo; split_bf16(v, hi, lo); dh[i * cop + c] = hi; dl[i * cop + c] = lo; }
            }
        for (; c < cop; c++) {
            if (fmt == SKIT_FMT_F32) dst[i * cop + c] = 0.f;
            else { dh[i * cop + c] = __float2bfloat16_rn(0.f); dl[i * cop + c] = __float2bfloat16_rn(0.f); }
        }
    }
}

__global__ void __launch_bounds__(256) operand_grad_to_nchw_kernel(const float* __restrict__ dpad, int n, int h, int w, int c, int pad, int pad_mode,
                                                                   int c0, int cs, float* dst, int accumulate) {
    const int hp = h + 2 * pad, wp = w + 2 * pad;
    const long long total = (long long)n * cs * h * w;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % w);
        long long t = i / w;
        const int y = (int)(t % h); t /= h;
        const int j = (int)(t % cs);
        const int b = (int)(t / cs);
        float v = 0.f;
        int ys[3], xs[3];
        int ny = 1, nx = 1;
        ys[0] = y + pad; xs[0] = x + pad;
        if (pad_mode == SKIT_PAD_REFLECT) {
            if (y >= 1 && y <= pad) ys[ny++] = pad - y;
            if (y <= h - 2 && y >= h - 1 - pad) ys[ny++] = pad + 2 * (h - 1) - y;
            if (x >= 1 && x <= pad) xs[nx++] = pad - x;
            if (x <= w - 2 && x >= w - 1 - pad) xs[nx++] = pad + 2 * (w - 1) - x;
        }
        for (int a = 0; a < ny; a++)
            for (int d = 0; d < nx; d++) v += dpad[(((long long)b * hp + ys[a]) * wp + xs[d]) * c + c0 + j];
        dst[i] = accumulate ? dst[i] + v : v;
    }
}

// ------------------------------------------------------------------------------ generator head
__global__ void __launch_bounds__(256) g_head_fwd_kernel(const float* __restrict__ raw, const float* __restrict__ mask, int n, int h, int w,
                                                         float scale_nz, float* fI, float* fT, float* fN) {
    const long long hw = (long long)h * w, total = (long long)n * hw;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / hw, pix = i - b * hw;
        const float m = mask ? mask[i] : 1.f;
        float t[5];
#pragma unroll
        for (int j = 0; j < 5; j++) t[j] = tanhf(raw[i * 5 + j]) * m;
#pragma unroll
        for (int j = 0; j < 3; j++) fI[(b * 3 + j) * hw + pix] = t[j];
        fT[(b * 2 + 0) * hw + pix] = t[3];
        fT[(b * 2 + 1) * hw + pix] = t[4];
        if (fN) {
            const float nrm = fmaxf(sqrtf(t[3] * t[3] + t[4] * t[4] + scale_nz * scale_nz), 1e-12f);
            fN[(b * 3 + 0) * hw + pix] = t[3] / nrm;
            fN[(b * 3 + 1) * hw + pix] = t[4] / nrm;
            fN[(b * 3 + 2) * hw + pix] = scale_nz / nrm;
        }
    }
}

__global__ void __launch_bounds__(256) g_head_bwd_kernel(const float* __restrict__ raw, const float* __restrict__ mask,
                                                         const float* __restrict__ dI, const float* __restrict__ dT,
                                                         int n, int h, int w, float* dst, __nv_bfloat16* dh, __nv_bfloat16* dl,
                                                         int cop, int fmt, int pad) {
    const int hp = h + 2 * pad, wp = w + 2 * pad;
    const long long hw = (long long)h * w, total = (long long)n * hp * wp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int px = (int)(i % wp);
        long long t = i / wp;
        const int py = (int)(t % hp);
        const long long b = t / hp;
        const int y = py - pad, x = px - pad;
        float v[5] = {0, 0, 0, 0, 0};
        if (y >= 0 && y < h && x >= 0 && x < w) {
            const long long pix = (long long)y * w + x;
            const float m = mask ? mask[b * hw + pix] : 1.f;
#pragma unroll
            for (int j = 0; j < 5; j++) {
                const float th = tanhf(raw[(b * hw + pix) * 5 + j]);
                const float d = j < 3 ? (dI ? dI[(b * 3 + j) * hw + pix] : 0.f) : (dT ? dT[(b * 2 + (j - 3)) * hw + pix] : 0.f);
                v[j] = d * m * (1.f - th * th);
            }
        }
        for (int j = 0; j < cop; j++) {   // cop >= 5: zero channel padding for the tensor-core path
            const float x = j < 5 ? v[j] : 0.f;
            if (fmt == SKIT_FMT_F32) dst[i * cop + j] = x;
            else { __nv_bfloat16 hi, lo; split_bf16(x, hi, lo); dh[i * cop + j] = hi; dl[i * cop + j] = lo; }
        }
    }
}

// Same gradient, written as two zero-haloed operands: the RGB (3 ch) and the touch (2 ch) decoders of the default
// U-Net generator end in separate transposed convs (networks.py:1635-1644).
__global__ void __launch_bounds__(256) g_head_bwd_split_kernel(const float* __restrict__ raw, const float* __restrict__ mask,
                                                               const float* __restrict__ dI, const float* __restrict__ dT,
                                                               int n, int h, int w, float* dstI, float* dstT, int pad) {
    const int hp = h + 2 * pad, wp = w + 2 * pad;
    const long long hw = (long long)h * w, total = (long long)n * hp * wp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int px = (int)(i % wp);
        long long t = i / wp;
        const int py = (int)(t % hp);
        const long long b = t / hp;
        const int y = py - pad, x = px - pad;
        float v[5] = {0, 0, 0, 0, 0};
        if (y >= 0 && y < h && x >= 0 && x < w) {
            const long long pix = (long long)y * w + x;
            const float m = mask ? mask[b * hw + pix] : 1.f;
#pragma unroll
            for (int j = 0; j < 5; j++) {
                const float th = tanhf(raw[(b * hw + pix) * 5 + j]);
                const float d = j < 3 ? (dI ? dI[(b * 3 + j) * hw + pix] : 0.f) : (dT ? dT[(b * 2 + (j - 3)) * hw + pix] : 0.f);
                v[j] = d * m * (1.f - th * th);
            }
        }
        dstI[i * 3 + 0] = v[0]; dstI[i * 3 + 1] = v[1]; dstI[i * 3 + 2] = v[2];
        dstT[i * 2 + 0] = v[3]; dstT[i * 2 + 1] = v[4];
    }
}

// x[n][c][hw] *= m[n][0][hw]  (set_input's background / contact masking, sinskitG_model.py:724,734,789-790, done on the
// device after the H2D copy of the raw tensors instead of on the host before it)
__global__ void __launch_bounds__(256) mask_mul_kernel(float* __restrict__ x, const float* __restrict__ m, int n, int c, long long hw) {
    const long long total = (long long)n * c * hw;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long pix = i % hw, b = i / (hw * c);
        x[i] *= m[b * hw + pix];
    }
}

// Channel mean of an NCHW image and its adjoint (the 1-channel "sketch-like" view of the generated RGB image that the
// PatchNCE query branch feeds back through the generator's encoder).
__global__ void __launch_bounds__(256) channel_mean_kernel(const float* __restrict__ x, int n, int c, long long hw, float* __restrict__ y) {
    const long long total = (long long)n * hw;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / hw, pix = i - b * hw;
        float s = 0.f;
        for (int j = 0; j < c; j++) s += x[(b * c + j) * hw + pix];
        y[i] = s / (float)c;
    }
}
__global__ void __launch_bounds__(256) channel_mean_bwd_kernel(const float* __restrict__ dy, int n, int c, long long hw, float* __restrict__ dx) {
    const long long total = (long long)n * c * hw;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long pix = i % hw, b = i / (hw * c);
        dx[i] += dy[b * hw + pix] / (float)c;
    }
}

// ------------------------------------------------------------------------------ DiffAugment 'bs' * M
__global__ void __launch_bounds__(256) diffaug_kernel(const float* __restrict__ x, const float* __restrict__ mask,
                                                      const float* __restrict__ ub, const float* __restrict__ us, int n, int h, int w, float* y) {
    const long long hw = (long long)h * w, total = (long long)n * hw;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long b = i / hw, pix = i - b * hw;
        const float bb = ub[b] - 0.5f, ss = us[b] * 2.f, m = mask[i];
        float v[3];
#pragma unroll
        for (int j = 0; j < 3; j++) v[j] = x[(b * 3 + j) * hw + pix] + bb;
        const float mean = (v[0] + v[1] + v[2]) / 3.f;
#pragma unroll
        for (int j = 0; j < 3; j++) y[(b * 3 + j) * hw + pix] = ((v[j] - mean) * ss + mean) * m;
    }
}

// ------------------------------------------------------------------------------ avg-pool pyramid
__global__ void __launch_bounds__(256) avgpool_fwd_kernel(const float* __restrict__ x, int planes, int h, int w, float* y) {
    const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
    const long long total = (long long)planes * ho * wo;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ox = (int)(i % wo);
        long long t = i / wo;
        const int oy = (int)(t % ho);
        const long long pl = t / ho;
        float s = 0.f; int cnt = 0;
        for (int a = 0; a < 3; a++) {
            const int sy = 2 * oy - 1 + a;
            if (sy < 0 || sy >= h) continue;
            for (int d = 0; d < 3; d++) {
                const int sx = 2 * ox - 1 + d;
                if (sx < 0 || sx >= w) continue;
                s += x[(pl * h + sy) * w + sx]; cnt++;
            }
        }
        y[i] = s / (float)cnt;
    }
}

__device__ inline int pool_valid(int o, int len) {  // number of in-range taps of output o on one axis
    int c = 0;
    for (int a = 0; a < 3; a++) { int s = 2 * o - 1 + a; c += (s >= 0 && s < len); }
    return c;
}

__global__ void __launch_bounds__(256) avgpool_bwd_kernel(const float* __restrict__ dy, int planes, int h, int w, float* dx, int accumulate) {
    const int ho = (h + 2 - 3) / 2 + 1, wo = (w + 2 - 3) / 2 + 1;
    const long long total = (long long)planes * h * w;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % w);
        long long t = i / w;
        const int y = (int)(t % h);
        const long long pl = t / h;
        float s = 0.f;
        for (int a = 0; a < 3; a++) {
            const int ty = y + 1 - a;
            if (ty < 0 || (ty & 1)) continue;
            const int oy = ty >> 1;
            if (oy >= ho) continue;
            const int cy = pool_valid(oy, h);
            for (int d = 0; d < 3; d++) {
                const int tx = x + 1 - d;
                if (tx < 0 || (tx & 1)) continue;
                const int ox = tx >> 1;
                if (ox >= wo) continue;
                s += dy[(pl * ho + oy) * wo + ox] / (float)(cy * pool_valid(ox, w));
            }
        }
        dx[i] = accumulate ? dx[i] + s : s;
    }
}

// ------------------------------------------------------------------------------ patch gather / scatter
struct GatherP {
    Srcs4 s;
    int h, w;
    const int* ox; const int* oy;
    int np, ps, ctot;
    float* dst;
};

// one thread per (patch, y, x); loops over the concatenated channels: reads are coalesced runs of
// ps floats per row, writes are fully coalesced.
__global__ void __launch_bounds__(256) patch_gather_kernel(GatherP g) {
    const long long total = (long long)g.np * g.ps * g.ps;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % g.ps);
        long long t = i / g.ps;
        const int y = (int)(t % g.ps);
        const int p = (int)(t / g.ps);
        const int sy = min(max(g.oy[p] + y, 0), g.h - 1), sx = min(max(g.ox[p] + x, 0), g.w - 1);
        for (int k = 0; k < g.s.n; k++)
            for (int j = 0; j < g.s.ch[k]; j++)
                g.dst[(((long long)p * g.ctot + g.s.coff[k] + j) * g.ps + y) * g.ps + x] =
                    __ldg(g.s.p[k] + ((long long)j * g.h + sy) * g.w + sx);
    }
}

__global__ void __launch_bounds__(256) patch_scatter_add_kernel(const float* __restrict__ dpatch, int ctot, int coff, int cs, int h, int w,
                                                                const int* ox, const int* oy, int np, int ps, float* dsrc) {
    const long long total = (long long)np * cs * ps * ps;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int x = (int)(i % ps);
        long long t = i / ps;
        const int y = (int)(t % ps); t /= ps;
        const int j = (int)(t % cs);
        const int p = (int)(t / cs);
        const int sy = min(max(oy[p] + y, 0), h - 1), sx = min(max(ox[p] + x, 0), w - 1);
        atomicAdd(dsrc + ((long long)j * h + sy) * w + sx, dpatch[(((long long)p * ctot + coff + j) * ps + y) * ps + x]);
    }
}

// ------------------------------------------------------------------------------ losses
__device__ inline float softplus_f(float x) { return x > 20.f ? x : log1pf(expf(x)); }

// grid: (chunks, n).  loss[b] += mean softplus(sign*pred)
__global__ void __launch_bounds__(256) gan_softplus_kernel(const float* __restrict__ pred, int hw, float sign, float* loss, float* dpred, float gscale) {
    __shared__ float red[8];
    const int b = blockIdx.y;
    const float inv = 1.f / (float)hw;
    float s = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
        const float z = sign * pred[(long long)b * hw + i];
        s += softplus_f(z);
        if (dpred) dpred[(long long)b * hw + i] = gscale * sign * inv / (1.f + expf(-z));
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int k = 0; k < 8; k++) tot += red[k];
        atomicAdd(loss + b, tot * inv);
    }
}

// GANLoss in every mode the reference defines (models/networks.py:500-522).  mode: 0 nonsaturating softplus(s x), 1 hinge relu(1 + s x),
// 2 wgan s x, 3 lsgan (x - t)^2, 4 vanilla BCE-with-logits(x, t); s = -1 for a real target, +1 for a fake one, t = the label.
// loss[b] += mean over the sample's hw elements; dpred = gscale * d(mean)/dx.
__global__ void __launch_bounds__(256) gan_loss_kernel(const float* __restrict__ pred, int hw, int mode, float sign, float target, float* loss,
                                                       float* dpred, float gscale) {
    __shared__ float red[8];
    const int b = blockIdx.y;
    const float inv = 1.f / (float)hw;
    float s = 0.f;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < hw; i += gridDim.x * blockDim.x) {
        const float x = pred[(long long)b * hw + i];
        float f, d;
        if (mode == 0) { const float z = sign * x; f = softplus_f(z); d = sign / (1.f + expf(-z)); }
        else if (mode == 1) { const float z = 1.f + sign * x; f = fmaxf(z, 0.f); d = z > 0.f ? sign : 0.f; }
        else if (mode == 2) { f = sign * x; d = sign; }
        else if (mode == 3) { const float e = x - target; f = e * e; d = 2.f * e; }
        else { f = fmaxf(x, 0.f) - x * target + log1pf(expf(-fabsf(x))); d = 1.f / (1.f + expf(-x)) - target; }
        s += f;
        if (dpred) dpred[(long long)b * hw + i] = gscale * inv * d;
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0) {
        float tot = 0.f;
        for (int k = 0; k < 8; k++) tot += red[k];
        atomicAdd(loss + b, tot * inv);
    }
}

__global__ void __launch_bounds__(256) l1_loss_kernel(const float* __restrict__ a, const float* __restrict__ b, long long numel, float scale,
                                                      float* loss, float* grad, float gscale, int accumulate) {
    __shared__ float red[8];
    float s = 0.f;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel; i += (long long)gridDim.x * blockDim.x) {
        const float d = a[i] - b[i];
        s += fabsf(d);
        if (grad) {
            const float gg = d > 0.f ? gscale : (d < 0.f ? -gscale : 0.f);
            grad[i] = accumulate ? grad[i] + gg : gg;
        }
    }
    s = warp_sum(s);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
    __syncthreads();
    if (threadIdx.x == 0 && loss) {
        float tot = 0.f;
        for (int k = 0; k < 8; k++) tot += red[k];
        atomicAdd(loss, tot * scale);
    }
}

// ------------------------------------------------------------------------------ Adam
__global__ void __launch_bounds__(256) adam_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                   long long numel, float lr_over_bc1, float inv_sqrt_bc2, float beta1, float beta2, float eps, float gs) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel; i += (long long)gridDim.x * blockDim.x) {
        const float gg = g[i] * gs;
        const float mm = beta1 * m[i] + (1.f - beta1) * gg;
        const float vv = beta2 * v[i] + (1.f - beta2) * gg * gg;
        m[i] = mm; v[i] = vv;
        p[i] -= lr_over_bc1 * mm / (sqrtf(vv) * inv_sqrt_bc2 + eps);
    }
}

// Same update with the step-dependent scalars read from device memory ([lr/bc1, 1/sqrt(bc2)]), so that a
// captured CUDA graph of the whole train step stays valid while the step count and learning rate advance.
__global__ void __launch_bounds__(256) adam_dev_kernel(float* __restrict__ p, const float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                                                       long long numel, const float* __restrict__ hyper, float beta1, float beta2, float eps, float gs) {
    const float lr_over_bc1 = hyper[0], inv_sqrt_bc2 = hyper[1];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < numel; i += (long long)gridDim.x * blockDim.x) {
        const float gg = g[i] * gs;
        const float mm = beta1 * m[i] + (1.f - beta1) * gg;
        const float vv = beta2 * v[i] + (1.f - beta2) * gg * gg;
        m[i] = mm; v[i] = vv;
        p[i] -= lr_over_bc1 * mm / (sqrtf(vv) * inv_sqrt_bc2 + eps);
    }
}

// ------------------------------------------------------------------------------ PatchNCE
// one warp per sampled row: coalesced gather of c floats, warp-shuffle L2 norm.
__global__ void __launch_bounds__(256) patch_sample_l2norm_kernel(const float* __restrict__ feat, int b, int hw, int c, const int* ids, int np,
                                                                  float* out, float* pre) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= b * np) return;
    const int bi = warp / np, pi = warp - bi * np;
    const float* src = feat + ((long long)bi * hw + ids[pi]) * c;
    float ss = 0.f;
    for (int j = lane; j < c; j += 32) { const float v = src[j]; ss += v * v; }
    ss = warp_sum(ss);
    const float inv = 1.f / (sqrtf(ss) + 1e-7f);
    for (int j = lane; j < c; j += 32) {
        const float v = src[j];
        out[(long long)warp * c + j] = v * inv;
        if (pre) pre[(long long)warp * c + j] = v;
    }
}

// dfeat[b][ids[i]][c] += drows[b*np + i][c]   (adjoint of the plain row gather of PatchSampleF, networks.py:706)
__global__ void __launch_bounds__(256) rows_scatter_add_kernel(const float* __restrict__ drows, int b, int hw, int c, const int* ids, int np, float* dfeat) {
    const long long total = (long long)b * np * c;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int ch = (int)(i % c);
        const long long row = i / c;
        const int bi = (int)(row / np), pi = (int)(row - (long long)bi * np);
        atomicAdd(dfeat + ((long long)bi * hw + ids[pi]) * c + ch, drows[i]);
    }
}

// y = x/(|x|+e): dx = dy/(|x|+e) - x * (x.dy) / (|x| (|x|+e)^2)
__global__ void __launch_bounds__(256) patch_sample_l2norm_bwd_kernel(const float* __restrict__ dout, const float* __restrict__ pre, int b, int hw, int c,
                                                                      const int* ids, int np, float* dfeat) {
    const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
    if (warp >= b * np) return;
    const int bi = warp / np, pi = warp - bi * np;
    const float* x = pre + (long long)warp * c;
    const float* dy = dout + (long long)warp * c;
    float ss = 0.f, dot = 0.f;
    for (int j = lane; j < c; j += 32) { ss += x[j] * x[j]; dot += x[j] * dy[j]; }
    ss = warp_sum(ss); dot = warp_sum(dot);
    const float nrm = sqrtf(ss), den = nrm + 1e-7f;
    const float k = nrm > 0.f ? dot / (nrm * den * den) : 0.f;
    float* dst = dfeat + ((long long)bi * hw + ids[pi]) * c;
    for (int j = lane; j < c; j += 32) atomicAdd(dst + j, dy[j] / den - x[j] * k);
}

// one warp per query row i of batch bi: logits[0] = q.k_i / T, logits[1+j] = (j==i ? -10 : q.k_j) / T
// loss = logsumexp(logits) - logits[0]; dq = sum_j softmax_j * dlogit_j/dq  (k detached).
// Two passes over the np keys (max, then sum) with warp-shuffle dot products; dim <= 1024.
__global__ void __launch_bounds__(256) patchnce_kernel(const float* __restrict__ q, const float* __restrict__ k, int b, int np, int dim, float inv_T,
                                                       float* loss, float* dq, float gscale) {
    extern __shared__ float sm[];  // per warp: logits[np + 1]
    const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int row = blockIdx.x * (blockDim.x >> 5) + wib;
    if (row >= b * np) return;
    const int bi = row / np, i = row - bi * np;
    float* lg = sm + (size_t)wib * (np + 1);
    const float* qi = q + (long long)row * dim;
    const float* kb = k + (long long)bi * np * dim;
    float mx = -INFINITY;
    for (int j = 0; j <= np; j++) {
        const float* kj = (j == 0) ? (k + (long long)row * dim) : (kb + (long long)(j - 1) * dim);
        float d = 0.f;
        for (int t = lane; t < dim; t += 32) d += qi[t] * kj[t];
        d = warp_sum(d);
        if (j > 0 && j - 1 == i) d = -10.f;
        d *= inv_T;
        if (lane == 0) lg[j] = d;
        mx = fmaxf(mx, d);
    }
    __syncwarp();
    float se = 0.f;
    for (int j = lane; j <= np; j += 32) se += expf(lg[j] - mx);
    se = warp_sum(se);
    const float lse = mx + logf(se);
    if (lane == 0) loss[row] = lse - lg[0];
    if (dq) {
        for (int t = lane; t < dim; t += 32) {
            float acc = 0.f;
            for (int j = 0; j <= np; j++) {
                if (j > 0 && j - 1 == i) continue;  // masked diagonal is a constant
                const float* kj = (j == 0) ? (k + (long long)row * dim) : (kb + (long long)(j - 1) * dim);
                float pj = expf(lg[j] - lse);
                if (j == 0) pj -= 1.f;
                acc += pj * kj[t];
            }
            dq[(long long)row * dim + t] = acc * inv_T * gscale;
        }
    }
}

// Tiled single-pass PatchNCE for dim <= 256 (every use on the hot path: feature widths 9 / 128 / 256 and the MLP width):
// a block owns RB query rows of one batch element (one warp per row); the keys stream through shared memory in tiles of 32
// (row stride dim + 1: lane j reads key j conflict-free), each lane owns one key of the tile for the logits and the columns
// t = lane + 32 u of dq, and the softmax is accumulated online (running max m, running sum s, rescaled accumulator) so K is
// read once.  loss = logsumexp(logits) - logit_pos;  dq = (sum_j p_j k_j - k_pos) / T * gscale  with the masked diagonal
// counted in the normaliser only (it is the constant -10 / T).
template <int RB>
__global__ void __launch_bounds__(RB * 32) patchnce_tiled_kernel(const float* __restrict__ q, const float* __restrict__ k, int b, int np, int dim,
                                                                 float inv_T, float* __restrict__ loss, float* __restrict__ dq, float gscale) {
    extern __shared__ float sm[];
    const int ld = dim + 1;
    float* ks = sm;                       // [32][ld]
    float* qs = ks + 32 * ld;             // [RB][dim]
    float* ps = qs + RB * dim;            // [RB][32]
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int tiles_per_b = (np + RB - 1) / RB;
    const int bi = blockIdx.x / tiles_per_b, i = (blockIdx.x - bi * tiles_per_b) * RB + warp;
    const bool valid = i < np;
    const long long row = (long long)bi * np + (valid ? i : 0);
    const float* kb = k + (long long)bi * np * dim;
    const float* kpos = k + row * dim;
    for (int t = lane; t < dim; t += 32) qs[warp * dim + t] = q[row * dim + t];
    __syncwarp();
    float l0 = 0.f;
    for (int t = lane; t < dim; t += 32) l0 += qs[warp * dim + t] * kpos[t];
    l0 = warp_sum(l0) * inv_T;
    float m = l0, ssum = 1.f, acc[8];
#pragma unroll
    for (int u = 0; u < 8; u++) { const int t = lane + 32 * u; acc[u] = t < dim ? kpos[t] : 0.f; }
    for (int j0 = 0; j0 < np; j0 += 32) {
        __syncthreads();
        const int nk = min(32, np - j0);
        for (int e = threadIdx.x; e < nk * dim; e += RB * 32) { const int jj = e / dim, t = e - jj * dim; ks[jj * ld + t] = kb[(long long)(j0 + jj) * dim + t]; }
        __syncthreads();
        const int j = j0 + lane;
        float lg = -INFINITY;
        if (lane < nk) {
            float d = 0.f;
            const float* kr = ks + lane * ld;
            const float* qr = qs + warp * dim;
            for (int t = 0; t < dim; t++) d += qr[t] * kr[t];
            lg = (j == i ? -10.f : d) * inv_T;
        }
        float tm = lg;
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) tm = fmaxf(tm, __shfl_xor_sync(0xffffffffu, tm, o));
        const float mn = fmaxf(m, tm), sc = __expf(m - mn);
        const float p = lane < nk ? __expf(lg - mn) : 0.f;
        ssum = ssum * sc + warp_sum(p);
        m = mn;
        ps[warp * 32 + lane] = (j == i) ? 0.f : p;
        __syncwarp();
#pragma unroll
        for (int u = 0; u < 8; u++) acc[u] *= sc;
        for (int jj = 0; jj < nk; jj++) {
            const float pj = ps[warp * 32 + jj];
            const float* kr = ks + jj * ld;
#pragma unroll
            for (int u = 0; u < 8; u++) { const int t = lane + 32 * u; if (t < dim) acc[u] += pj * kr[t]; }
        }
    }
    if (!valid) return;
    if (lane == 0) loss[row] = m + logf(ssum) - l0;
    if (dq) {
        const float inv_s = 1.f / ssum;
#pragma unroll
        for (int u = 0; u < 8; u++) { const int t = lane + 32 * u; if (t < dim) dq[row * dim + t] = (acc[u] * inv_s - kpos[t]) * inv_T * gscale; }
    }
}

}  // namespace skit

using namespace skit;

extern "C" const char* skit_last_error(void) { return g_err; }
extern "C" int skit_version(void) { return 1; }
extern "C" int skit_built_arch(void) { return 100; }

static int fill_srcs(Srcs4& s, const float* const* srcs, const int* chans, const int* coffs, int nsrc) {
    if (!srcs || !chans || nsrc < 1 || nsrc > 4) return -1;
    s.n = nsrc;
    int off = 0;
    for (int i = 0; i < nsrc; i++) {
        if (!srcs[i] || chans[i] < 1) return -1;
        s.p[i] = srcs[i]; s.ch[i] = chans[i]; s.coff[i] = coffs ? coffs[i] : off;
        off += chans[i];
    }
    return off;
}

/* Zero-fill as a memset node (cudaMemsetAsync): inside a captured step graph this replaces the fill kernels torch.zeros / zero_()
 * would launch for the gradient buckets, the split-K scratch of the weight gradients and the scatter targets. */
extern "C" int skit_zero_bytes(void* p, long long nbytes, void* stream) {
    SKIT_REQUIRE(p && nbytes >= 0, "zero_bytes: bad arguments");
    if (nbytes == 0) return SKIT_OK;
    cudaError_t e = cudaMemsetAsync(p, 0, (size_t)nbytes, as_stream(stream));
    if (e != cudaSuccess) {
        set_error("cudaMemsetAsync failed: %s", cudaGetErrorString(e));
        return SKIT_ERR_CUDA;
    }
    return SKIT_OK;
}

extern "C" int skit_nchw_cat_to_operand(const float* const* srcs, const int* chans, int nsrc,
                                        int n, int h, int w, const skit_operand* op, int pad, int pad_mode, void* stream) {
    Srcs4 s{};
    int ctot = fill_srcs(s, srcs, chans, nullptr, nsrc);
    SKIT_REQUIRE(ctot > 0 && op && op->p0 && (op->fmt == SKIT_FMT_F32 || (op->fmt == SKIT_FMT_BF16X2 && op->p1)), "nchw_cat_to_operand: bad sources or operand");
    SKIT_REQUIRE(op->n == n && op->hp == h + 2 * pad && op->wp == w + 2 * pad && op->c >= ctot,
                 "nchw_cat_to_operand: operand dims [%d,%d,%d,%d] != [%d,%d,%d,%d]", op->n, op->hp, op->wp, op->c, n, h + 2 * pad, w + 2 * pad, ctot);
    SKIT_REQUIRE(pad_mode != SKIT_PAD_REFLECT || (pad < h && pad < w), "nchw_cat_to_operand: reflect pad too large");
    const long long total = (long long)n * (h + 2 * pad) * (w + 2 * pad);
    nchw_cat_to_operand_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(s, n, h, w, ctot, op->c, op->fmt, (float*)op->p0,
                                                                                    (__nv_bfloat16*)op->p0, (__nv_bfloat16*)op->p1, pad, pad_mode);
    return check_launch("nchw_cat_to_operand_kernel");
}

extern "C" int skit_operand_grad_to_nchw(const float* dpad, int n, int h, int w, int c, int pad, int pad_mode,
                                         int c0, int cs, float* dst, int accumulate, void* stream) {
    SKIT_REQUIRE(dpad && dst && n > 0 && h > 0 && w > 0 && c0 >= 0 && cs > 0 && c0 + cs <= c, "operand_grad_to_nchw: bad arguments");
    SKIT_REQUIRE(pad_mode == SKIT_PAD_ZERO || pad_mode == SKIT_PAD_REFLECT, "operand_grad_to_nchw: unsupported pad mode");
    const long long total = (long long)n * cs * h * w;
    operand_grad_to_nchw_kernel<<<grid_for(total, 256), 256, 0, as_stream(stream)>>>(dpad, n, h, w, c, pad, pad_mode, c0, cs, dst, accumulate);
    return check_launch("operand_grad_to_nchw_kernel");
}

extern "C" int skit_g_head_fwd(const float* raw, const float* mask, int n, int h, int w, float scale_nz,
                               float* fake_I, float* fake_T, float* fake_N, void* stream) {
    SKIT_REQUIRE(raw && fake_I && fake_T && n > 0 && h > 0 && w > 0, "g_head_fwd: bad arguments");
    g_head_fwd_kernel<<<grid_for((long long)n * h * w, 256), 256, 0, as_stream(stream)>>>(raw, mask, n, h, w, scale_nz, fake_I, fake_T, fake_N);
    return check_launch("g_head_fwd_kernel");
}

extern "C" int skit_g_head_bwd(const float* raw, const float* mask, const float* dI, const float* dT,
                               int n, int h, int w, const skit_operand* op, int pad, void* stream) {
    SKIT_REQUIRE(raw && (dI || dT) && op && op->p0 && (op->fmt == SKIT_FMT_F32 || (op->fmt == SKIT_FMT_BF16X2 && op->p1)), "g_head_bwd: bad arguments");
    SKIT_REQUIRE(op->n == n && op->hp == h + 2 * pad && op->wp == w + 2 * pad && op->c >= 5, "g_head_bwd: operand dims mismatch");
    g_head_bwd_kernel<<<grid_for((long long)n * op->hp * op->wp, 256), 256, 0, as_stream(stream)>>>(
        raw, mask, dI, dT, n, h, w, (float*)op->p0, (__nv_bfloat16*)op->p0, (__nv_bfloat16*)op->p1, op->c, op->fmt, pad);
    return check_launch("g_head_bwd_kernel");
}

extern "C" int skit_g_head_bwd_split(const float* raw, const float* mask, const float* dI, const float* dT,
                                     int n, int h, int w, const skit_operand* opI, const skit_operand* opT, int pad, void* stream) {
    SKIT_REQUIRE(raw && (dI || dT) && opI && opT && opI->p0 && opT->p0 && opI->fmt == SKIT_FMT_F32 && opT->fmt == SKIT_FMT_F32,
                 "g_head_bwd_split: bad arguments (fp32 operands required)");
    SKIT_REQUIRE(opI->n == n && opI->hp == h + 2 * pad && opI->wp == w + 2 * pad && opI->c == 3 &&
                 opT->n == n && opT->hp == h + 2 * pad && opT->wp == w + 2 * pad && opT->c == 2, "g_head_bwd_split: operand dims mismatch");
    g_head_bwd_split_kernel<<<grid_for((long long)n * opI->hp * opI->wp, 256), 256, 0, as_stream(stream)>>>(
        raw, mask, dI, dT, n, h, w, (float*)opI->p0, (float*)opT->p0, pad);
    return check_launch("g_head_bwd_split_kernel");
}

extern "C" int skit_mask_mul(float* x, const float* m, int n, int c, int h, int w, void* stream) {
    SKIT_REQUIRE(x && m && n > 0 && c > 0 && h > 0 && w > 0, "mask_mul: bad arguments");
    mask_mul_kernel<<<grid_for((long long)n * c * h * w, 256), 256, 0, as_stream(stream)>>>(x, m, n, c, (long long)h * w);
    return check_launch("mask_mul_kernel");
}

extern "C" int skit_channel_mean(const float* x, int n, int c, int h, int w, float* y, void* stream) {
    SKIT_REQUIRE(x && y && n > 0 && c > 0 && h > 0 && w > 0, "channel_mean: bad arguments");
    channel_mean_kernel<<<grid_for((long long)n * h * w, 256), 256, 0, as_stream(stream)>>>(x, n, c, (long long)h * w, y);
    return check_launch("channel_mean_kernel");
}

extern "C" int skit_channel_mean_bwd(const float* dy, int n, int c, int h, int w, float* dx, void* stream) {
    SKIT_REQUIRE(dy && dx && n > 0 && c > 0 && h > 0 && w > 0, "channel_mean_bwd: bad arguments");
    channel_mean_bwd_kernel<<<grid_for((long long)n * c * h * w, 256), 256, 0, as_stream(stream)>>>(dy, n, c, (long long)h * w, dx);
    return check_launch("channel_mean_bwd_kernel");
}

extern "C" int skit_diffaug_bs_mask(const float* x, const float* mask, const float* u_b, const float* u_s,
                                    int n, int h, int w, float* y, void* stream) {
    SKIT_REQUIRE(x && mask && u_b && u_s && y && n > 0 && h > 0 && w > 0, "diffaug_bs_mask: bad arguments");
    diffaug_kernel<<<grid_for((long long)n * h * w, 256), 256, 0, as_stream(stream)>>>(x, mask, u_b, u_s, n, h, w, y);
    return check_launch("diffaug_kernel");
}

extern "C" int skit_avgpool3s2_fwd(const float* x, int planes, int h, int w, float* y, void* stream) {
    SKIT_REQUIRE(x && y && planes > 0 && h > 0 && w > 0, "avgpool3s2_fwd: bad arguments");
    const int ho = (h - 1) / 2 + 1, wo = (w - 1) / 2 + 1;
    avgpool_fwd_kernel<<<grid_for((long long)planes * ho * wo, 256), 256, 0, as_stream(stream)>>>(x, planes, h, w, y);
    return check_launch("avgpool_fwd_kernel");
}

extern "C" int skit_avgpool3s2_bwd(const float* dy, int planes, int h, int w, float* dx, int accumulate, void* stream) {
    SKIT_REQUIRE(dy && dx && planes > 0 && h > 0 && w > 0, "avgpool3s2_bwd: bad arguments");
    avgpool_bwd_kernel<<<grid_for((long long)planes * h * w, 256), 256, 0, as_stream(stream)>>>(dy, planes, h, w, dx, accumulate);
    return check_launch("avgpool_bwd_kernel");
}

extern "C" int skit_patch_gather(const float* const* srcs, const int* chans, const int* coffs, int nsrc,
                                 int h, int w, const int* ox, const int* oy, int np, int ps,
                                 float* dst, int ctot, void* stream) {
    GatherP g{};
    int tot = fill_srcs(g.s, srcs, chans, coffs, nsrc);
    SKIT_REQUIRE(tot > 0 && ox && oy && dst && np > 0 && ps > 0 && h > 0 && w > 0, "patch_gather: bad arguments");
    for (int i = 0; i < nsrc; i++)
        SKIT_REQUIRE(g.s.coff[i] >= 0 && g.s.coff[i] + g.s.ch[i] <= ctot, "patch_gather: source %d does not fit in %d channels", i, ctot);
    g.h = h; g.w = w; g.ox = ox; g.oy = oy; g.np = np; g.ps = ps; g.ctot = ctot; g.dst = dst;
    patch_gather_kernel<<<grid_for((long long)np * ps * ps, 256), 256, 0, as_stream(stream)>>>(g);
    return check_launch("patch_gather_kernel");
}

extern "C" int skit_patch_scatter_add(const float* dpatch, int ctot, int coff, int cs, int h, int w,
                                      const int* ox, const int* oy, int np, int ps, float* dsrc, void* stream) {
    SKIT_REQUIRE(dpatch && dsrc && ox && oy && np > 0 && ps > 0 && coff >= 0 && cs > 0 && coff + cs <= ctot, "patch_scatter_add: bad arguments");
    patch_scatter_add_kernel<<<grid_for((long long)np * cs * ps * ps, 256), 256, 0, as_stream(stream)>>>(dpatch, ctot, coff, cs, h, w, ox, oy, np, ps, dsrc);
    return check_launch("patch_scatter_add_kernel");
}

extern "C" int skit_gan_softplus(const float* pred, int n, int hw, float sign, float* loss, float* dpred, float gscale, void* stream) {
    SKIT_REQUIRE(pred && loss && n > 0 && hw > 0, "gan_softplus: bad arguments");
    dim3 grid(min(cdiv(hw, 256), 64), n);
    gan_softplus_kernel<<<grid, 256, 0, as_stream(stream)>>>(pred, hw, sign, loss, dpred, gscale);
    return check_launch("gan_softplus_kernel");
}

extern "C" int skit_gan_loss(const float* pred, int n, int hw, int mode, int target_is_real, float target, float* loss, float* dpred,
                             float gscale, void* stream) {
    SKIT_REQUIRE(pred && loss && n > 0 && hw > 0 && mode >= 0 && mode <= 4, "gan_loss: bad arguments (mode %d)", mode);
    dim3 grid(min(cdiv(hw, 256), 64), n);
    gan_loss_kernel<<<grid, 256, 0, as_stream(stream)>>>(pred, hw, mode, target_is_real ? -1.f : 1.f, target, loss, dpred, gscale);
    return check_launch("gan_loss_kernel");
}

extern "C" int skit_l1_loss(const float* a, const float* b, long long numel, float scale, float* loss,
                            float* grad, float gscale, int accumulate, void* stream) {
    SKIT_REQUIRE(a && b && numel > 0 && (loss || grad), "l1_loss: bad arguments");
    l1_loss_kernel<<<grid_for(numel, 256, 4), 256, 0, as_stream(stream)>>>(a, b, numel, scale, loss, grad, gscale, accumulate);
    return check_launch("l1_loss_kernel");
}

extern "C" int skit_adam_step(float* p, const float* g, float* m, float* v, long long numel, int step,
                              float lr, float beta1, float beta2, float eps, float grad_scale, void* stream) {
    SKIT_REQUIRE(p && g && m && v && numel > 0 && step >= 1, "adam_step: bad arguments");
    const double bc1 = 1.0 - pow((double)beta1, step), bc2 = 1.0 - pow((double)beta2, step);
    adam_kernel<<<grid_for(numel, 256), 256, 0, as_stream(stream)>>>(p, g, m, v, numel, (float)(lr / bc1), (float)(1.0 / sqrt(bc2)), beta1, beta2, eps, grad_scale);
    return check_launch("adam_kernel");
}

extern "C" int skit_adam_step_dev(float* p, const float* g, float* m, float* v, long long numel, const float* hyper,
                                  float beta1, float beta2, float eps, float grad_scale, void* stream) {
    SKIT_REQUIRE(p && g && m && v && hyper && numel > 0, "adam_step_dev: bad arguments");
    adam_dev_kernel<<<grid_for(numel, 256), 256, 0, as_stream(stream)>>>(p, g, m, v, numel, hyper, beta1, beta2, eps, grad_scale);
    return check_launch("adam_dev_kernel");
}

extern "C" int skit_patch_sample_l2norm(const float* feat, int b, int hw, int c, const int* ids, int np,
                                        float* out, float* pre, void* stream) {
    SKIT_REQUIRE(feat && ids && out && b > 0 && hw > 0 && c > 0 && np > 0, "patch_sample_l2norm: bad arguments");
    patch_sample_l2norm_kernel<<<cdiv(b * np, 8), 256, 0, as_stream(stream)>>>(feat, b, hw, c, ids, np, out, pre);
    return check_launch("patch_sample_l2norm_kernel");
}

extern "C" int skit_patch_sample_l2norm_bwd(const float* dout, const float* pre, int b, int hw, int c,
                                            const int* ids, int np, float* dfeat, void* stream) {
    SKIT_REQUIRE(dout && pre && ids && dfeat && b > 0 && hw > 0 && c > 0 && np > 0, "patch_sample_l2norm_bwd: bad arguments");
    patch_sample_l2norm_bwd_kernel<<<cdiv(b * np, 8), 256, 0, as_stream(stream)>>>(dout, pre, b, hw, c, ids, np, dfeat);
    return check_launch("patch_sample_l2norm_bwd_kernel");
}

extern "C" int skit_rows_scatter_add(const float* drows, int b, int hw, int c, const int* ids, int np, float* dfeat, void* stream) {
    SKIT_REQUIRE(drows && ids && dfeat && b > 0 && hw > 0 && c > 0 && np > 0, "rows_scatter_add: bad arguments");
    rows_scatter_add_kernel<<<grid_for((long long)b * np * c, 256), 256, 0, as_stream(stream)>>>(drows, b, hw, c, ids, np, dfeat);
    return check_launch("rows_scatter_add_kernel");
}

extern "C" int skit_patchnce(const float* q, const float* k, int b, int np, int dim, float inv_T,
                             float* loss, float* dq, float gscale, void* stream) {
    SKIT_REQUIRE(q && k && loss && b > 0 && np > 0 && dim > 0, "patchnce: bad arguments");
    if (dim <= 256) {
        constexpr int RB = 4;
        const size_t sm2 = ((size_t)32 * (dim + 1) + (size_t)RB * dim + RB * 32) * sizeof(float);
        patchnce_tiled_kernel<RB><<<b * cdiv(np, RB), RB * 32, sm2, as_stream(stream)>>>(q, k, b, np, dim, inv_T, loss, dq, gscale);
        return check_launch("patchnce_tiled_kernel");
    }
    const size_t smem = (size_t)8 * (np + 1) * sizeof(float);
    SKIT_REQUIRE(smem <= 48 * 1024, "patchnce: num_patches %d too large", np);
    patchnce_kernel<<<cdiv(b * np, 8), 256, smem, as_stream(stream)>>>(q, k, b, np, dim, inv_T, loss, dq, gscale);
    return check_launch("patchnce_kernel");
}
