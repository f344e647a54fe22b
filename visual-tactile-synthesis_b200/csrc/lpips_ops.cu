// LPIPS-VGG16 perceptual loss support (pip `lpips` 0.1.4, `LPIPS(net='vgg')`, called at models/sinskitG_model.py:495,
// 1639-1645, 1711).  The VGG16 trunk itself is thirteen 3x3 zero-padded convs + ReLU and runs on the library's conv kernels
// (skit_conv2d_fwd / skit_conv2d_dgrad_s1 with frozen packs); this file holds what sits around them:
//   * ScalingLayer -> the stem conv's haloed operand (and its transpose back to an NCHW image gradient),
//   * MaxPool2d(2, 2) forward into the next conv's operand, and backward (first maximum of the window, like ATen),
//   * one LPIPS layer: unit-normalise both feature maps over channels, squared difference, 1x1 `lin` weights, spatial mean
//     — value and the gradient w.r.t. the first feature map in the same pass.
#include "skit_common.cuh"

namespace skit {

__constant__ float kLpipsShift[3] = {-0.030f, -0.088f, -0.188f};
__constant__ float kLpipsScale[3] = {0.458f, 0.448f, 0.450f};

// x: NCHW [n][cin][h][w], cin = 1 (broadcast to the three channels, as ScalingLayer's (inp - shift) / scale does) or 3.
// op: fp32 operand [n][h+2][w+2][3], zero halo.
__global__ void lpips_scale_fwd_kernel(const float* __restrict__ x, int n, int cin, int h, int w, float* __restrict__ op) {
    const int hp = h + 2, wp = w + 2;
    const long long total = (long long)n * hp * wp;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int xp = (int)(i % wp); long long t = i / wp;
        const int yp = (int)(t % hp); const int b = (int)(t / hp);
        const int y = yp - 1, xx = xp - 1;
        float v[3] = {0.f, 0.f, 0.f};
        if (y >= 0 && y < h && xx >= 0 && xx < w) {
#pragma unroll
            for (int c = 0; c < 3; c++) {
                const float s = x[(((long long)b * cin + (cin == 1 ? 0 : c)) * h + y) * w + xx];
                v[c] = (s - kLpipsShift[c]) / kLpipsScale[c];
            }
        }
        op[i * 3] = v[0]; op[i * 3 + 1] = v[1]; op[i * 3 + 2] = v[2];
    }
}

// dop: gradient w.r.t. the haloed operand [n][h+2][w+2][3]; dx[n][cin][h][w] (+)= gscale * dop / scale (summed over the
// three channels for a 1-channel input).  dx_ctot / dx_c0: dx is channels [dx_c0, dx_c0 + cin) of an [n][dx_ctot][h][w] tensor.
__global__ void lpips_scale_bwd_kernel(const float* __restrict__ dop, int n, int cin, int h, int w, float gscale,
                                       float* __restrict__ dx, int dx_ctot, int dx_c0, int accumulate) {
    const long long total = (long long)n * h * w;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int xx = (int)(i % w); long long t = i / w;
        const int y = (int)(t % h); const int b = (int)(t / h);
        const float* g = dop + (((long long)b * (h + 2) + y + 1) * (w + 2) + xx + 1) * 3;
        float d[3];
#pragma unroll
        for (int c = 0; c < 3; c++) d[c] = gscale * g[c] / kLpipsScale[c];
        if (cin == 1) {
            float* o = dx + (((long long)b * dx_ctot + dx_c0) * h + y) * w + xx;
            const float s = d[0] + d[1] + d[2];
            *o = accumulate ? *o + s : s;
        } else {
#pragma unroll
            for (int c = 0; c < 3; c++) {
                float* o = dx + (((long long)b * dx_ctot + dx_c0 + c) * h + y) * w + xx;
                *o = accumulate ? *o + d[c] : d[c];
            }
        }
    }
}

// f: dense NHWC [n][h][w][c] (h, w even) -> haloed operand [n][h/2+2p][w/2+2p][c], fp32 or bf16 hi/lo, zero halo.
__global__ void __launch_bounds__(256) maxpool2_fwd_kernel(const float* __restrict__ f, int n, int h, int w, int c, int pad, int fmt,
                                                           float* o0, __nv_bfloat16* oh, __nv_bfloat16* ol) {
    const int ho = h / 2, wo = w / 2, hp = ho + 2 * pad, wp = wo + 2 * pad, cv = c >> 2;
    const long long total = (long long)n * hp * wp * cv;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int q = (int)(i % cv); long long t = i / cv;
        const int xp = (int)(t % wp); t /= wp;
        const int yp = (int)(t % hp); const int b = (int)(t / hp);
        const int y = yp - pad, x = xp - pad;
        float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
        if (y >= 0 && y < ho && x >= 0 && x < wo) {
            const float* s = f + (((long long)b * h + 2 * y) * w + 2 * x) * c + q * 4;
            const float4 a = *reinterpret_cast<const float4*>(s), bb = *reinterpret_cast<const float4*>(s + c);
            const float4 cc = *reinterpret_cast<const float4*>(s + (long long)w * c), d = *reinterpret_cast<const float4*>(s + (long long)w * c + c);
            v.x = fmaxf(fmaxf(a.x, bb.x), fmaxf(cc.x, d.x)); v.y = fmaxf(fmaxf(a.y, bb.y), fmaxf(cc.y, d.y));
            v.z = fmaxf(fmaxf(a.z, bb.z), fmaxf(cc.z, d.z)); v.w = fmaxf(fmaxf(a.w, bb.w), fmaxf(cc.w, d.w));
        }
        const long long dst = i * 4;
        if (fmt == SKIT_FMT_F32) *reinterpret_cast<float4*>(o0 + dst) = v;
        else {
            const float vv[4] = {v.x, v.y, v.z, v.w};
            __align__(8) __nv_bfloat16 hh[4];
            __align__(8) __nv_bfloat16 ll[4];
#pragma unroll
            for (int j = 0; j < 4; j++) split_bf16(vv[j], hh[j], ll[j]);
            *reinterpret_cast<uint2*>(oh + dst) = *reinterpret_cast<uint2*>(hh);
            *reinterpret_cast<uint2*>(ol + dst) = *reinterpret_cast<uint2*>(ll);
        }
    }
}

// df[n][h][w][c] = (add ? add : 0) + dpool routed to the FIRST maximum of each 2x2 window (row-major scan, ATen's choice).
// dpool: gradient w.r.t. the pooled haloed operand [n][h/2+2p][w/2+2p][c] (fp32).
__global__ void __launch_bounds__(256) maxpool2_bwd_kernel(const float* __restrict__ f, const float* __restrict__ dpool, int n, int h, int w,
                                                           int c, int pad, const float* __restrict__ add, float* __restrict__ df) {
    const int ho = h / 2, wo = w / 2, hp = ho + 2 * pad, wp = wo + 2 * pad, cv = c >> 2;
    const long long total = (long long)n * ho * wo * cv;
    for (long long i = blockIdx.x * 256LL + threadIdx.x; i < total; i += (long long)gridDim.x * 256) {
        const int q = (int)(i % cv); long long t = i / cv;
        const int x = (int)(t % wo); t /= wo;
        const int y = (int)(t % ho); const int b = (int)(t / ho);
        const long long s00 = (((long long)b * h + 2 * y) * w + 2 * x) * c + q * 4;
        const long long offs[4] = {s00, s00 + c, s00 + (long long)w * c, s00 + (long long)w * c + c};
        float4 in[4];
#pragma unroll
        for (int k = 0; k < 4; k++) in[k] = *reinterpret_cast<const float4*>(f + offs[k]);
        const float4 g = *reinterpret_cast<const float4*>(dpool + (((long long)b * hp + y + pad) * wp + x + pad) * c + q * 4);
        const float gg[4] = {g.x, g.y, g.z, g.w};
        float out[4][4];
#pragma unroll
        for (int j = 0; j < 4; j++) {
            const float v[4] = {(&in[0].x)[j], (&in[1].x)[j], (&in[2].x)[j], (&in[3].x)[j]};
            int best = 0;
#pragma unroll
            for (int k = 1; k < 4; k++) if (v[k] > v[best]) best = k;
#pragma unroll
            for (int k = 0; k < 4; k++) out[k][j] = (k == best) ? gg[j] : 0.f;
        }
#pragma unroll
        for (int k = 0; k < 4; k++) {
            float4 o = make_float4(out[k][0], out[k][1], out[k][2], out[k][3]);
            if (add) { const float4 a = *reinterpret_cast<const float4*>(add + offs[k]); o.x += a.x; o.y += a.y; o.z += a.z; o.w += a.w; }
            *reinterpret_cast<float4*>(df + offs[k]) = o;
        }
    }
}

// One LPIPS layer (lpips/lpips.py: normalize_tensor, (f0-f1)**2, NetLinLayer 1x1 conv, spatial_average).
// f0, f1: [n][h][w][c] post-ReLU features.  loss[b] += sum_pixels sum_c lw[c] * (u0 - u1)^2 / (h w), u = f / (|f| + 1e-10).
// df0 (optional) = gscale * d loss[b] / d f0.  One warp per pixel.
template <int CMAX4>   // float4 per lane held in registers (c <= 128 * CMAX4)
__global__ void __launch_bounds__(256) lpips_layer_kernel(const float* __restrict__ f0, const float* __restrict__ f1, const float* __restrict__ lw,
                                                          int n, int hw, int c, float gscale, float* __restrict__ loss, float* __restrict__ df0) {
    __shared__ float red[8];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const int b = blockIdx.y;
    const int c4 = c >> 2;
    float acc = 0.f;
    for (int p = blockIdx.x * 8 + warp; p < hw; p += gridDim.x * 8) {
        const long long base = ((long long)b * hw + p) * c;
        float4 a[CMAX4], q[CMAX4];
        float s0 = 0.f, s1 = 0.f;
#pragma unroll
        for (int k = 0; k < CMAX4; k++) {
            const int idx = lane + 32 * k;
            a[k] = make_float4(0.f, 0.f, 0.f, 0.f); q[k] = a[k];
            if (idx < c4) {
                a[k] = *reinterpret_cast<const float4*>(f0 + base + idx * 4);
                q[k] = *reinterpret_cast<const float4*>(f1 + base + idx * 4);
            }
            s0 += a[k].x * a[k].x + a[k].y * a[k].y + a[k].z * a[k].z + a[k].w * a[k].w;
            s1 += q[k].x * q[k].x + q[k].y * q[k].y + q[k].z * q[k].z + q[k].w * q[k].w;
        }
        s0 = warp_sum(s0); s1 = warp_sum(s1);
        const float r0 = sqrtf(s0), r1 = sqrtf(s1);
        const float i0 = 1.f / (r0 + 1e-10f), i1 = 1.f / (r1 + 1e-10f);
        float val = 0.f, gf = 0.f;     // gf = sum_c g_c * f0_c with g_c = 2 lw_c (u0_c - u1_c)
        float4 g[CMAX4];
#pragma unroll
        for (int k = 0; k < CMAX4; k++) {
            const int idx = lane + 32 * k;
            g[k] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (idx < c4) {
                const float4 wv = *reinterpret_cast<const float4*>(lw + idx * 4);
                const float d0 = a[k].x * i0 - q[k].x * i1, d1 = a[k].y * i0 - q[k].y * i1;
                const float d2 = a[k].z * i0 - q[k].z * i1, d3 = a[k].w * i0 - q[k].w * i1;
                val += wv.x * d0 * d0 + wv.y * d1 * d1 + wv.z * d2 * d2 + wv.w * d3 * d3;
                g[k] = make_float4(2.f * wv.x * d0, 2.f * wv.y * d1, 2.f * wv.z * d2, 2.f * wv.w * d3);
                gf += g[k].x * a[k].x + g[k].y * a[k].y + g[k].z * a[k].z + g[k].w * a[k].w;
            }
        }
        acc += val;
        if (df0) {
            gf = warp_sum(gf);
            // d u_j / d f_c = delta_jc / (r + eps) - f_j f_c / (r (r + eps)^2); an all-zero pixel (r = 0) gets the first term only
            const float k2 = r0 > 0.f ? gf * i0 * i0 / r0 : 0.f;
            const float sc = gscale / (float)hw;
#pragma unroll
            for (int k = 0; k < CMAX4; k++) {
                const int idx = lane + 32 * k;
                if (idx < c4) {
                    float4 o;
                    o.x = sc * (g[k].x * i0 - a[k].x * k2); o.y = sc * (g[k].y * i0 - a[k].y * k2);
                    o.z = sc * (g[k].z * i0 - a[k].z * k2); o.w = sc * (g[k].w * i0 - a[k].w * k2);
                    *reinterpret_cast<float4*>(df0 + base + idx * 4) = o;
                }
            }
        }
    }
    acc = warp_sum(acc);
    if (lane == 0) red[warp] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        float t = 0.f;
        for (int i = 0; i < 8; i++) t += red[i];
        atomicAdd(loss + b, t / (float)hw);
    }
}

}  // namespace skit

using namespace skit;

static int grid_for(long long total) { return (int)min((long long)148 * 16, cdivll(total, 256)); }

extern "C" int skit_lpips_scale_fwd(const float* x, int n, int cin, int h, int w, const skit_operand* op, void* stream) {
    SKIT_REQUIRE(x && op && op->p0 && n > 0 && h > 0 && w > 0 && (cin == 1 || cin == 3), "lpips_scale_fwd: bad arguments (1 or 3 input channels)");
    SKIT_REQUIRE(op->fmt == SKIT_FMT_F32 && op->n == n && op->c == 3 && op->hp == h + 2 && op->wp == w + 2,
                 "lpips_scale_fwd: needs an fp32 [n][h+2][w+2][3] operand");
    lpips_scale_fwd_kernel<<<grid_for((long long)n * (h + 2) * (w + 2)), 256, 0, as_stream(stream)>>>(x, n, cin, h, w, (float*)op->p0);
    return check_launch("lpips_scale_fwd_kernel");
}

extern "C" int skit_lpips_scale_bwd(const float* dop, int n, int cin, int h, int w, float gscale, float* dx, int dx_ctot, int dx_c0,
                                    int accumulate, void* stream) {
    SKIT_REQUIRE(dop && dx && n > 0 && h > 0 && w > 0 && (cin == 1 || cin == 3) && dx_c0 >= 0 && dx_c0 + cin <= dx_ctot,
                 "lpips_scale_bwd: bad arguments");
    lpips_scale_bwd_kernel<<<grid_for((long long)n * h * w), 256, 0, as_stream(stream)>>>(dop, n, cin, h, w, gscale, dx, dx_ctot, dx_c0, accumulate);
    return check_launch("lpips_scale_bwd_kernel");
}

extern "C" int skit_maxpool2_fwd(const float* f, int n, int h, int w, int c, const skit_operand* op, int pad, void* stream) {
    SKIT_REQUIRE(f && op && op->p0 && n > 0 && h > 0 && w > 0 && h % 2 == 0 && w % 2 == 0 && c % 4 == 0 && pad >= 0,
                 "maxpool2_fwd: bad arguments (even sizes, channels a multiple of 4)");
    SKIT_REQUIRE(op->n == n && op->c == c && op->hp == h / 2 + 2 * pad && op->wp == w / 2 + 2 * pad, "maxpool2_fwd: operand dims mismatch");
    SKIT_REQUIRE(op->fmt == SKIT_FMT_F32 || op->p1, "maxpool2_fwd: bf16x2 operand without its lo plane");
    const long long total = (long long)n * op->hp * op->wp * (c / 4);
    maxpool2_fwd_kernel<<<grid_for(total), 256, 0, as_stream(stream)>>>(f, n, h, w, c, pad, op->fmt, (float*)op->p0,
                                                                        (__nv_bfloat16*)op->p0, (__nv_bfloat16*)op->p1);
    return check_launch("maxpool2_fwd_kernel");
}

extern "C" int skit_maxpool2_bwd(const float* f, const float* dpool, int n, int h, int w, int c, int pad, const float* add, float* df,
                                 void* stream) {
    SKIT_REQUIRE(f && dpool && df && n > 0 && h % 2 == 0 && w % 2 == 0 && h > 0 && w > 0 && c % 4 == 0 && pad >= 0, "maxpool2_bwd: bad arguments");
    maxpool2_bwd_kernel<<<grid_for((long long)n * (h / 2) * (w / 2) * (c / 4)), 256, 0, as_stream(stream)>>>(f, dpool, n, h, w, c, pad, add, df);
    return check_launch("maxpool2_bwd_kernel");
}

extern "C" int skit_lpips_layer(const float* f0, const float* f1, const float* lin_w, int n, int h, int w, int c, float gscale,
                                float* loss, float* df0, void* stream) {
    SKIT_REQUIRE(f0 && f1 && lin_w && loss && n > 0 && h > 0 && w > 0 && c > 0 && c % 4 == 0 && c <= 512,
                 "lpips_layer: bad arguments (channels a multiple of 4, at most 512)");
    const int hw = h * w;
    dim3 grid(min(cdiv(hw, 8), 148 * 8 / max(1, min(n, 8)) + 1), n);
    if (c <= 128) lpips_layer_kernel<1><<<grid, 256, 0, as_stream(stream)>>>(f0, f1, lin_w, n, hw, c, gscale, loss, df0);
    else if (c <= 256) lpips_layer_kernel<2><<<grid, 256, 0, as_stream(stream)>>>(f0, f1, lin_w, n, hw, c, gscale, loss, df0);
    else lpips_layer_kernel<4><<<grid, 256, 0, as_stream(stream)>>>(f0, f1, lin_w, n, hw, c, gscale, loss, df0);
    return check_launch("lpips_layer_kernel");
}
