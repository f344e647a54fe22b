// StyleGAN2 generator support (reference: models/stylegan_networks.py).
//
// The reference runs every resampling layer as two passes: a [1,3,3,1] x [1,3,3,1] upfirdn blur (F.pad + F.conv2d per channel,
// :38-72) and a strided / transposed convolution.  Both are linear with zero padding, so the blur is folded into the
// filter on the device each time the weights change and the layer becomes ONE ordinary zero-padded convolution that
// the tcgen05 conv kernels run directly:
//   blur(pad 2,2) -> conv3x3 stride 2     ==  conv6x6 stride 2 pad 2               (ConvLayer downsample, :625-643)
//   blur(pad 1,1) -> conv1x1 stride 2     ==  conv4x4 stride 2 pad 1               (ResBlock skip, :677)
//   conv_transpose3x3 stride 2 -> blur(pad 1,1, x4)  ==  four 3x3 pad-1 convs, one per output parity, stacked along the
//                                                      output channels (phase-major) + depth-to-space (ModulatedConv2d :323-334)
// The element-wise half (bias, noise, leaky-ReLU * sqrt 2, residual merge, depth-to-space, halo write, bf16 split for the
// next conv's operand) is one pass (sg2_bias_act_kernel).
#include "skit_common.cuh"

namespace skit {

__device__ __forceinline__ float blur_tap(int j) { return (j == 0 || j == 3) ? 0.125f : 0.375f; }   // [1,3,3,1] / 8

// One block per output channel o.  mode 0: equalised weight (scale only).  mode 1: down composite ((k+3)^2 taps).
// mode 2: demodulated up composite, 4 phases x 3x3.
__global__ void __launch_bounds__(256) sg2_weight_prep_kernel(const float* __restrict__ w, int co, int ci, int k, int mode,
                                                              float* __restrict__ out) {
    __shared__ float red[8];
    __shared__ float s_demod;
    const int o = blockIdx.x;
    const int kk = k * k;
    const float scale = rsqrtf((float)(ci * kk));
    const float* wo = w + (long long)o * ci * kk;
    if (mode == 2) {   // demodulation: rsqrt(sum (scale*w)^2 + 1e-8) over (ci, k, k)  (stylegan_networks.py:313-315)
        float s = 0.f;
        for (int i = threadIdx.x; i < ci * kk; i += 256) { const float v = scale * wo[i]; s = fmaf(v, v, s); }
        s = warp_sum(s);
        if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = s;
        __syncthreads();
        if (threadIdx.x == 0) {
            float t = 0.f;
            for (int i = 0; i < 8; i++) t += red[i];
            s_demod = rsqrtf(t + 1e-8f);
        }
        __syncthreads();
    }
    if (mode == 0) {
        for (int i = threadIdx.x; i < ci * kk; i += 256) out[(long long)o * ci * kk + i] = scale * wo[i];
        return;
    }
    if (mode == 1) {
        const int K = k + 3;
        for (int i = threadIdx.x; i < ci * K * K; i += 256) {
            const int c = i / (K * K), r = i - c * K * K, u = r / K, v = r - u * K;
            float acc = 0.f;
            for (int t = 0; t < k; t++) {
                const int j = u - t;
                if (j < 0 || j > 3) continue;
                for (int s = 0; s < k; s++) {
                    const int jj = v - s;
                    if (jj < 0 || jj > 3) continue;
                    acc = fmaf(scale * wo[(c * k + t) * k + s], blur_tap(j) * blur_tap(jj), acc);
                }
            }
            out[((long long)o * ci + c) * K * K + r] = acc;
        }
        return;
    }
    // mode 2: k == 3.  c2[sy][sx] = sum w[t][s] * b4[t - sy + 1] * b4[s - sx + 1], b4 = [1,3,3,1]/4;
    // phase (py,px) tap (a,b) = c2[2 + py - 2a][2 + px - 2b]
    const float dm = s_demod * scale;
    for (int i = threadIdx.x; i < ci * 36; i += 256) {
        const int c = i / 36, r = i - c * 36, ph = r / 9, tap = r - ph * 9;
        const int py = ph >> 1, px = ph & 1, a = tap / 3, b = tap - a * 3;
        const int sy = 2 + py - 2 * a, sx = 2 + px - 2 * b;
        float acc = 0.f;
        for (int t = 0; t < 3; t++) {
            const int jy = t - sy + 1;
            if (jy < 0 || jy > 3) continue;
            for (int s = 0; s < 3; s++) {
                const int jx = s - sx + 1;
                if (jx < 0 || jx > 3) continue;
                acc = fmaf(dm * wo[(c * 3 + t) * 3 + s], 4.f * blur_tap(jy) * blur_tap(jx), acc);
            }
        }
        out[(((long long)ph * co + o) * ci + c) * 9 + tap] = acc;
    }
}

struct Sg2ActP {
    const float* raw; int n, h, w, craw, c;     // raw: [n][h][w][craw]; h, w: RAW resolution
    const float* bias;                          // [c] or null
    const float* noise; const float* noise_w;   // noise [n][H][W] at output resolution, strength = *noise_w
    const float* skip;                          // dense [n][H][W][c] or null
    int shuffle;                                // depth-to-space 2x: raw channels (py*2+px)*c + ch -> pixel (2y+py, 2x+px)
    int act;                                    // 1: leaky-ReLU(0.2) * gain, 0: identity
    float gain, post;
    float* dense;                               // [n][H][W][c] or null
    float* o0; __nv_bfloat16* oh; __nv_bfloat16* ol; int fmt, pad;   // haloed operand [n][H+2p][W+2p][c] or null
    float* nchw; int nchw_c;                    // [n][nchw_c][H][W] or null (nchw_c <= c: drops zero-padded channels)
};

// out = (lrelu(raw + bias + nw * noise) * gain + skip) * post.  One thread per 4 channels of a padded output pixel.
__global__ void __launch_bounds__(256) sg2_bias_act_kernel(Sg2ActP p) {
    const int H = p.shuffle ? 2 * p.h : p.h, W = p.shuffle ? 2 * p.w : p.w;
    const int hp = H + 2 * p.pad, wp = W + 2 * p.pad, cv = p.c >> 2;
    const bool has_op = p.o0 || p.oh;
    const long long total = (long long)p.n * hp * wp * cv;
    const float nw = p.noise ? __ldg(p.noise_w) : 0.f;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int q = (int)(i % cv); long long t = i / cv;
        const int xp = (int)(t % wp); t /= wp;
        const int yp = (int)(t % hp); const int n = (int)(t / hp);
        const int ch = q * 4, y = yp - p.pad, x = xp - p.pad;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        const bool inside = y >= 0 && y < H && x >= 0 && x < W;
        if (inside) {
            long long src;
            if (p.shuffle) src = (((long long)n * p.h + (y >> 1)) * p.w + (x >> 1)) * p.craw + ((y & 1) * 2 + (x & 1)) * p.c + ch;
            else src = (((long long)n * p.h + y) * p.w + x) * p.craw + ch;
            v = *reinterpret_cast<const float4*>(p.raw + src);
            if (p.bias) { v.x += __ldg(p.bias + ch); v.y += __ldg(p.bias + ch + 1); v.z += __ldg(p.bias + ch + 2); v.w += __ldg(p.bias + ch + 3); }   // scalar loads: a bias inside a flat parameter bucket is only 4-byte aligned
            if (p.noise) { const float nz = nw * p.noise[((long long)n * H + y) * W + x]; v.x += nz; v.y += nz; v.z += nz; v.w += nz; }
            if (p.act) {
                v.x = (v.x > 0.f ? v.x : 0.2f * v.x) * p.gain; v.y = (v.y > 0.f ? v.y : 0.2f * v.y) * p.gain;
                v.z = (v.z > 0.f ? v.z : 0.2f * v.z) * p.gain; v.w = (v.w > 0.f ? v.w : 0.2f * v.w) * p.gain;
            }
            const long long pix = ((long long)n * H + y) * W + x;
            if (p.skip) {
                const float4 s = *reinterpret_cast<const float4*>(p.skip + pix * p.c + ch);
                v.x += s.x; v.y += s.y; v.z += s.z; v.w += s.w;
            }
            v.x *= p.post; v.y *= p.post; v.z *= p.post; v.w *= p.post;
            if (p.dense) *reinterpret_cast<float4*>(p.dense + pix * p.c + ch) = v;
            if (p.nchw) {
                const long long plane = (long long)H * W, base = ((long long)n * p.nchw_c + ch) * plane + (long long)y * W + x;
                const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                for (int j = 0; j < 4; j++) if (ch + j < p.nchw_c) p.nchw[base + j * plane] = vv[j];
            }
        }
        if (has_op) {
            const long long dst = (((long long)n * hp + yp) * wp + xp) * p.c + ch;
            if (p.fmt == SKIT_FMT_F32) *reinterpret_cast<float4*>(p.o0 + dst) = v;
            else {
                __nv_bfloat16 h0, l0, h1, l1, h2, l2, h3, l3;
                split_bf16(v.x, h0, l0); split_bf16(v.y, h1, l1); split_bf16(v.z, h2, l2); split_bf16(v.w, h3, l3);
                __nv_bfloat162 a = __halves2bfloat162(h0, h1), b = __halves2bfloat162(h2, h3);
                __nv_bfloat162 c = __halves2bfloat162(l0, l1), d = __halves2bfloat162(l2, l3);
                uint2 hv, lv;
                hv.x = *reinterpret_cast<uint32_t*>(&a); hv.y = *reinterpret_cast<uint32_t*>(&b);
                lv.x = *reinterpret_cast<uint32_t*>(&c); lv.y = *reinterpret_cast<uint32_t*>(&d);
                *reinterpret_cast<uint2*>(p.oh + dst) = hv;
                *reinterpret_cast<uint2*>(p.ol + dst) = lv;
            }
        }
    }
}

// Transpose of sg2_weight_prep_kernel: dw[o] += (d eff / d w[o])^T deff.  One block per output channel; mode 2 stages the
// gradient w.r.t. the demodulated filter in shared memory (ci * 9 floats) for the demodulation's rank-1 correction
//   v = scale w, wd = v d, d = rsqrt(sum v^2 + eps):  dL/dv = d G - d^3 v sum(G v).
__global__ void __launch_bounds__(256) sg2_weight_prep_bwd_kernel(const float* __restrict__ w, int co, int ci, int k, int mode,
                                                                  const float* __restrict__ deff, float* __restrict__ dw) {
    extern __shared__ float sG[];
    __shared__ float red[2][8];
    __shared__ float s_tot[2];
    const int o = blockIdx.x;
    const int kk = k * k;
    const float scale = rsqrtf((float)(ci * kk));
    const float* wo = w + (long long)o * ci * kk;
    float* dwo = dw + (long long)o * ci * kk;
    if (mode == 0) {
        for (int i = threadIdx.x; i < ci * kk; i += 256) dwo[i] += scale * deff[(long long)o * ci * kk + i];
        return;
    }
    if (mode == 1) {
        const int K = k + 3;
        for (int i = threadIdx.x; i < ci * kk; i += 256) {
            const int c = i / kk, r = i - c * kk, t = r / k, s = r - t * k;
            const float* e = deff + ((long long)o * ci + c) * K * K;
            float acc = 0.f;
#pragma unroll
            for (int j = 0; j < 4; j++)
#pragma unroll
                for (int jj = 0; jj < 4; jj++) acc = fmaf(e[(t + j) * K + s + jj], blur_tap(j) * blur_tap(jj), acc);
            dwo[i] += scale * acc;
        }
        return;
    }
    // mode 2
    float sv2 = 0.f, sgv = 0.f;
    for (int i = threadIdx.x; i < ci * 9; i += 256) {
        const int c = i / 9, r = i - c * 9, t = r / 3, s = r - t * 3;
        float g = 0.f;
        for (int ph = 0; ph < 4; ph++) {
            const int py = ph >> 1, px = ph & 1;
            const float* e = deff + (((long long)ph * co + o) * ci + c) * 9;
            for (int a = 0; a < 3; a++) {
                const int jy = t - 1 - py + 2 * a;
                if (jy < 0 || jy > 3) continue;
                for (int b = 0; b < 3; b++) {
                    const int jx = s - 1 - px + 2 * b;
                    if (jx < 0 || jx > 3) continue;
                    g = fmaf(e[a * 3 + b], 4.f * blur_tap(jy) * blur_tap(jx), g);
                }
            }
        }
        sG[i] = g;
        const float v = scale * wo[i];
        sv2 = fmaf(v, v, sv2); sgv = fmaf(g, v, sgv);
    }
    sv2 = warp_sum(sv2); sgv = warp_sum(sgv);
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = sv2; red[1][threadIdx.x >> 5] = sgv; }
    __syncthreads();
    if (threadIdx.x == 0) {
        float a = 0.f, b = 0.f;
        for (int i = 0; i < 8; i++) { a += red[0][i]; b += red[1][i]; }
        s_tot[0] = rsqrtf(a + 1e-8f); s_tot[1] = b;
    }
    __syncthreads();
    const float d = s_tot[0], S = s_tot[1];
    for (int i = threadIdx.x; i < ci * 9; i += 256) {
        const float v = scale * wo[i];
        dwo[i] += scale * (d * sG[i] - d * d * d * v * S);
    }
}

struct Sg2BwdP {
    const float* raw; int n, h, w, craw, c;     // as in the forward pass (h, w: raw resolution)
    const float* bias; const float* noise; const float* noise_w;
    int shuffle, act; float gain, post;
    const float* dpa; int pa;                   // gradient w.r.t. the OUTPUT as haloed tensors [n][H+2p][W+2p][c] ...
    const float* dpb; int pb;
    const float* dd;                            // ... and / or dense [n][H][W][c]; d_out = sum of those given
    float* o0; __nv_bfloat16* oh; __nv_bfloat16* ol; int fmt, q;     // d_raw operand [n][h+2q][w+2q][craw_eff], zero halo
    float* dskip;                               // optional dense [n][H][W][c] = d_out * post (gradient of the skip input)
    float* dbias; float* dnoise_w;              // optional accumulators (+=)
};

// One thread per channel quad of a padded RAW-operand pixel.  d_raw = d_out * post * (act ? gain * lrelu'(t) : 1),
// t = raw + bias + nw * noise.
__global__ void __launch_bounds__(256) sg2_bias_act_bwd_kernel(Sg2BwdP p) {
    __shared__ float s_db[512];
    __shared__ float s_red[8];
    const int ce = p.shuffle ? 4 * p.c : p.c;            // channels of the raw gradient operand
    const int cv = ce >> 2, cq = p.c >> 2;
    const int H = p.shuffle ? 2 * p.h : p.h, W = p.shuffle ? 2 * p.w : p.w;
    const int hq = p.h + 2 * p.q, wq = p.w + 2 * p.q;
    const long long total = (long long)p.n * hq * wq * cv;
    for (int i = threadIdx.x; i < p.c; i += 256) s_db[i] = 0.f;
    __syncthreads();
    const float nw = p.noise ? __ldg(p.noise_w) : 0.f;
    float dn = 0.f;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int qd = (int)(i % cv); long long t = i / cv;
        const int xq = (int)(t % wq); t /= wq;
        const int yq = (int)(t % hq); const int n = (int)(t / hq);
        const int yr = yq - p.q, xr = xq - p.q;
        float4 g = make_float4(0.f, 0.f, 0.f, 0.f);
        if (yr >= 0 && yr < p.h && xr >= 0 && xr < p.w) {
            const int ph = p.shuffle ? qd / cq : 0, ch = (p.shuffle ? qd - ph * cq : qd) * 4;
            const int y = p.shuffle ? 2 * yr + (ph >> 1) : yr, x = p.shuffle ? 2 * xr + (ph & 1) : xr;
            const long long pix = ((long long)n * H + y) * W + x;
            if (p.dd) g = *reinterpret_cast<const float4*>(p.dd + pix * p.c + ch);
            if (p.dpa) {
                const float4 a = *reinterpret_cast<const float4*>(p.dpa + (((long long)n * (H + 2 * p.pa) + y + p.pa) * (W + 2 * p.pa) + x + p.pa) * p.c + ch);
                g.x += a.x; g.y += a.y; g.z += a.z; g.w += a.w;
            }
            if (p.dpb) {
                const float4 a = *reinterpret_cast<const float4*>(p.dpb + (((long long)n * (H + 2 * p.pb) + y + p.pb) * (W + 2 * p.pb) + x + p.pb) * p.c + ch);
                g.x += a.x; g.y += a.y; g.z += a.z; g.w += a.w;
            }
            g.x *= p.post; g.y *= p.post; g.z *= p.post; g.w *= p.post;
            if (p.dskip) *reinterpret_cast<float4*>(p.dskip + pix * p.c + ch) = g;
            if (p.act) {
                float4 v = *reinterpret_cast<const float4*>(p.raw + (((long long)n * p.h + yr) * p.w + xr) * p.craw + (p.shuffle ? ph * p.c : 0) + ch);
                if (p.bias) { v.x += __ldg(p.bias + ch); v.y += __ldg(p.bias + ch + 1); v.z += __ldg(p.bias + ch + 2); v.w += __ldg(p.bias + ch + 3); }   // scalar loads: a bias inside a flat parameter bucket is only 4-byte aligned
                float nz = 0.f;
                if (p.noise) { nz = p.noise[pix]; const float a = nw * nz; v.x += a; v.y += a; v.z += a; v.w += a; }
                g.x *= p.gain * (v.x > 0.f ? 1.f : 0.2f); g.y *= p.gain * (v.y > 0.f ? 1.f : 0.2f);
                g.z *= p.gain * (v.z > 0.f ? 1.f : 0.2f); g.w *= p.gain * (v.w > 0.f ? 1.f : 0.2f);
                if (p.noise) dn += nz * (g.x + g.y + g.z + g.w);
            }
            if (p.dbias) {
                atomicAdd(&s_db[ch], g.x); atomicAdd(&s_db[ch + 1], g.y); atomicAdd(&s_db[ch + 2], g.z); atomicAdd(&s_db[ch + 3], g.w);
            }
        }
        const long long dst = i * 4;
        if (p.fmt == SKIT_FMT_F32) *reinterpret_cast<float4*>(p.o0 + dst) = g;
        else {
            const float vv[4] = {g.x, g.y, g.z, g.w};
            __align__(8) __nv_bfloat16 hh[4];
            __align__(8) __nv_bfloat16 ll[4];
#pragma unroll
            for (int j = 0; j < 4; j++) split_bf16(vv[j], hh[j], ll[j]);
            *reinterpret_cast<uint2*>(p.oh + dst) = *reinterpret_cast<uint2*>(hh);
            *reinterpret_cast<uint2*>(p.ol + dst) = *reinterpret_cast<uint2*>(ll);
        }
    }
    if (p.dnoise_w) {
        dn = warp_sum(dn);
        if ((threadIdx.x & 31) == 0) s_red[threadIdx.x >> 5] = dn;
    }
    __syncthreads();
    if (p.dbias) for (int i = threadIdx.x; i < p.c; i += 256) if (s_db[i] != 0.f) atomicAdd(p.dbias + i, s_db[i]);
    if (p.dnoise_w && threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; i++) t += s_red[i];
        atomicAdd(p.dnoise_w, t);
    }
}

}  // namespace skit

using namespace skit;

extern "C" int skit_sg2_weight_prep(const float* w, int co, int ci, int k, int mode, float* out, void* stream) {
    SKIT_REQUIRE(w && out && co > 0 && ci > 0 && k > 0 && mode >= 0 && mode <= 2, "sg2_weight_prep: bad arguments");
    SKIT_REQUIRE(mode != 2 || k == 3, "sg2_weight_prep: the up-sampling composite is built for 3x3 filters (got k=%d)", k);
    sg2_weight_prep_kernel<<<co, 256, 0, as_stream(stream)>>>(w, co, ci, k, mode, out);
    return check_launch("sg2_weight_prep_kernel");
}

extern "C" int skit_sg2_bias_act(const float* raw, int n, int h, int w, int craw, int c, const float* bias,
                                 const float* noise, const float* noise_w, const float* skip, int shuffle,
                                 int act, float gain, float post, float* dense, const skit_operand* op, int pad, float* nchw,
                                 int nchw_c, void* stream) {
    SKIT_REQUIRE(raw && n > 0 && h > 0 && w > 0 && c > 0 && c % 4 == 0 && craw % 4 == 0, "sg2_bias_act: bad arguments (channels must be multiples of 4)");
    SKIT_REQUIRE(craw >= c * (shuffle ? 4 : 1), "sg2_bias_act: raw has %d channels, too few for c=%d", craw, c);
    SKIT_REQUIRE((noise == nullptr) == (noise_w == nullptr), "sg2_bias_act: noise and its strength go together");
    SKIT_REQUIRE(dense || op || nchw, "sg2_bias_act: no output");
    SKIT_REQUIRE(nchw == nullptr || (nchw_c > 0 && nchw_c <= c), "sg2_bias_act: nchw_c must be in [1, c]");
    const int H = shuffle ? 2 * h : h, W = shuffle ? 2 * w : w;
    Sg2ActP p{};
    p.raw = raw; p.n = n; p.h = h; p.w = w; p.craw = craw; p.c = c; p.bias = bias; p.noise = noise; p.noise_w = noise_w;
    p.skip = skip; p.shuffle = shuffle; p.act = act; p.gain = gain; p.post = post;
    p.dense = dense; p.nchw = nchw; p.nchw_c = nchw_c; p.pad = 0; p.fmt = SKIT_FMT_F32;
    if (op) {
        SKIT_REQUIRE(op->p0 && op->n == n && op->c == c && op->hp == H + 2 * pad && op->wp == W + 2 * pad && pad >= 0,
                     "sg2_bias_act: operand dims do not match the output (%dx%d pad %d)", H, W, pad);
        SKIT_REQUIRE(op->fmt == SKIT_FMT_F32 || op->p1, "sg2_bias_act: bf16x2 operand without its lo plane");
        p.pad = pad; p.fmt = op->fmt;
        if (op->fmt == SKIT_FMT_F32) p.o0 = (float*)op->p0;
        else { p.oh = (__nv_bfloat16*)op->p0; p.ol = (__nv_bfloat16*)op->p1; }
    }
    const long long total = (long long)n * (H + 2 * p.pad) * (W + 2 * p.pad) * (c / 4);
    const int blocks = (int)min((long long)148 * 16, cdivll(total, 256));
    sg2_bias_act_kernel<<<blocks, 256, 0, as_stream(stream)>>>(p);
    return check_launch("sg2_bias_act_kernel");
}

extern "C" int skit_sg2_weight_prep_bwd(const float* w, int co, int ci, int k, int mode, const float* deff, float* dw, void* stream) {
    SKIT_REQUIRE(w && deff && dw && co > 0 && ci > 0 && k > 0 && mode >= 0 && mode <= 2 && (mode != 2 || k == 3), "sg2_weight_prep_bwd: bad arguments");
    const size_t smem = mode == 2 ? (size_t)ci * 9 * sizeof(float) : 0;
    SKIT_REQUIRE(smem <= 48 * 1024, "sg2_weight_prep_bwd: ci = %d too wide for the shared-memory stage", ci);
    sg2_weight_prep_bwd_kernel<<<co, 256, smem, as_stream(stream)>>>(w, co, ci, k, mode, deff, dw);
    return check_launch("sg2_weight_prep_bwd_kernel");
}

extern "C" int skit_sg2_bias_act_bwd(const float* raw, int n, int h, int w, int craw, int c, const float* bias,
                                     const float* noise, const float* noise_w, int shuffle, int act, float gain, float post,
                                     const float* dpad_a, int pad_a, const float* dpad_b, int pad_b, const float* ddense,
                                     const skit_operand* draw, int q, float* dskip, float* dbias, float* dnoise_w, void* stream) {
    SKIT_REQUIRE(n > 0 && h > 0 && w > 0 && c > 0 && c % 4 == 0 && c <= 512 && draw && draw->p0, "sg2_bias_act_bwd: bad arguments (c a multiple of 4, at most 512)");
    SKIT_REQUIRE(dpad_a || dpad_b || ddense, "sg2_bias_act_bwd: no incoming gradient");
    SKIT_REQUIRE(!act || raw, "sg2_bias_act_bwd: the activation's backward needs the saved raw conv output");
    SKIT_REQUIRE((noise == nullptr) == (noise_w == nullptr), "sg2_bias_act_bwd: noise and its strength go together");
    SKIT_REQUIRE(!(shuffle && dskip), "sg2_bias_act_bwd: no skip on the up-sampling layer");
    const int ce = shuffle ? 4 * c : c;
    SKIT_REQUIRE(draw->n == n && draw->c == ce && draw->hp == h + 2 * q && draw->wp == w + 2 * q && q >= 0,
                 "sg2_bias_act_bwd: gradient operand dims mismatch (want [%d][%d][%d][%d])", n, h + 2 * q, w + 2 * q, ce);
    SKIT_REQUIRE(draw->fmt == SKIT_FMT_F32 || draw->p1, "sg2_bias_act_bwd: bf16x2 operand without its lo plane");
    Sg2BwdP p{};
    p.raw = raw; p.n = n; p.h = h; p.w = w; p.craw = craw; p.c = c; p.bias = bias; p.noise = noise; p.noise_w = noise_w;
    p.shuffle = shuffle; p.act = act; p.gain = gain; p.post = post;
    p.dpa = dpad_a; p.pa = pad_a; p.dpb = dpad_b; p.pb = pad_b; p.dd = ddense;
    p.fmt = draw->fmt; p.q = q;
    if (draw->fmt == SKIT_FMT_F32) p.o0 = (float*)draw->p0;
    else { p.oh = (__nv_bfloat16*)draw->p0; p.ol = (__nv_bfloat16*)draw->p1; }
    p.dskip = dskip; p.dbias = act ? dbias : nullptr; p.dnoise_w = (act && noise) ? dnoise_w : nullptr;
    const long long total = (long long)n * (h + 2 * q) * (w + 2 * q) * (ce / 4);
    const int blocks = (int)min((long long)148 * 8, cdivll(total, 256));
    sg2_bias_act_bwd_kernel<<<blocks, 256, 0, as_stream(stream)>>>(p);
    return check_launch("sg2_bias_act_bwd_kernel");
}
