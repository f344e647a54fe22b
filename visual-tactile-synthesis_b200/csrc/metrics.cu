// Evaluation metrics of compute_evaluation_metric (models/model_utils.py:431-561) as single-pass reduction kernels:
// the min-max rescale of the images, PSNR's squared error, SSIM (11x11 Gaussian windows, sigma 1.5), the surface-normal angle
// error of the touch patches (models/normal_losses.py:10-33) and their MSE.  All HBM-bound: one read of each input, fp64
// accumulators (a 1536 x 1536 x 3 image is 7 M terms per sum).  The reference runs these every print_freq iterations inside
// get_current_visuals on the full-resolution outputs; here they stay on the device and return device scalars.
#include "skit_common.cuh"

namespace skit {

__device__ __forceinline__ void atomic_min_f(float* addr, float v) {   // valid for any sign: ordered-int trick
    if (v >= 0.f) atomicMin(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMax(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}
__device__ __forceinline__ void atomic_max_f(float* addr, float v) {
    if (v >= 0.f) atomicMax(reinterpret_cast<int*>(addr), __float_as_int(v));
    else atomicMin(reinterpret_cast<unsigned int*>(addr), __float_as_uint(v));
}

// out[0] = min(x), out[1] = max(x); out must be initialised to {+inf, -inf}
__global__ void __launch_bounds__(256) minmax_kernel(const float* __restrict__ x, long long n, float* out) {
    float lo = INFINITY, hi = -INFINITY;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float v = x[i];
        lo = fminf(lo, v); hi = fmaxf(hi, v);
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o));
        hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o));
    }
    if ((threadIdx.x & 31) == 0) { atomic_min_f(out, lo); atomic_max_f(out + 1, hi); }
}

// sum over all elements of (a' - b')^2 with a' = (a - lo) * s, b' = clamp((b - lo) * s, 0, 1) when mm is given ({lo, hi} on the
// device: the min-max rescale of model_utils.py:483-487), or a' = a, b' = clamp(b, 0, 1) when clamp_b, else plain.
__global__ void __launch_bounds__(256) sq_err_kernel(const float* __restrict__ a, const float* __restrict__ b, long long n,
                                                     const float* __restrict__ mm, int clamp_b, double* out) {
    float lo = 0.f, sc = 1.f;
    if (mm) { lo = mm[0]; sc = 1.f / (mm[1] - mm[0]); }
    double acc = 0.0;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
        const float av = (a[i] - lo) * sc;
        float bv = (b[i] - lo) * sc;
        if (mm || clamp_b) bv = fminf(fmaxf(bv, 0.f), 1.f);
        const float d = av - bv;
        acc += (double)(d * d);
    }
    acc = warp_sum_d(acc);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

// SSIM with an 11-tap Gaussian window (sigma 1.5), reflect padding of 5, and the 5-pixel border of the index map dropped
// (torchmetrics functional/image/ssim.py).  One block = one 32 x 32 tile of one (image, channel) plane: the (32+10)^2 inputs of
// both images go to shared memory (rescaled / clamped as in sq_err_kernel), then a horizontal and a vertical pass over the five
// moment maps.  out[0] += sum of the index over the kept pixels.
constexpr int SS_T = 32, SS_R = 5, SS_H = SS_T + 2 * SS_R;
__global__ void __launch_bounds__(256) ssim_kernel(const float* __restrict__ a, const float* __restrict__ b, int h, int w,
                                                   const float* __restrict__ mm, float c1, float c2, double* out) {
    __shared__ float sa[SS_H][SS_H + 1], sb[SS_H][SS_H + 1];
    __shared__ float hm[5][SS_H][SS_T + 1];      // horizontally filtered moments: a, b, aa, bb, ab
    __shared__ float g[11];
    __shared__ double red[8];
    const long long plane = (long long)blockIdx.z * h * w;
    const int x0 = blockIdx.x * SS_T, y0 = blockIdx.y * SS_T;
    float lo = 0.f, sc = 1.f;
    if (mm) { lo = mm[0]; sc = 1.f / (mm[1] - mm[0]); }
    if (threadIdx.x < 11) {
        float s = 0.f;
        for (int i = 0; i < 11; i++) { const float d = (float)(i - 5); s += expf(-(d / 1.5f) * (d / 1.5f) * 0.5f); }
        const float d = (float)((int)threadIdx.x - 5);
        g[threadIdx.x] = expf(-(d / 1.5f) * (d / 1.5f) * 0.5f) / s;
    }
    for (int i = threadIdx.x; i < SS_H * SS_H; i += 256) {
        const int py = i / SS_H, px = i - py * SS_H;
        const int sy = pad_src(y0 + py, SS_R, h, SKIT_PAD_REFLECT), sx = pad_src(x0 + px, SS_R, w, SKIT_PAD_REFLECT);
        float av = 0.f, bv = 0.f;
        if (sy >= 0 && sy < h && sx >= 0 && sx < w) {
            av = (a[plane + (long long)sy * w + sx] - lo) * sc;
            bv = (b[plane + (long long)sy * w + sx] - lo) * sc;
            if (mm) bv = fminf(fmaxf(bv, 0.f), 1.f);
        }
        sa[py][px] = av; sb[py][px] = bv;
    }
    __syncthreads();
    for (int i = threadIdx.x; i < SS_H * SS_T; i += 256) {
        const int py = i / SS_T, ox = i - py * SS_T;
        float m0 = 0.f, m1 = 0.f, m2 = 0.f, m3 = 0.f, m4 = 0.f;
#pragma unroll
        for (int t = 0; t < 11; t++) {
            const float av = sa[py][ox + t], bv = sb[py][ox + t], wt = g[t];
            m0 += wt * av; m1 += wt * bv; m2 += wt * av * av; m3 += wt * bv * bv; m4 += wt * av * bv;
        }
        hm[0][py][ox] = m0; hm[1][py][ox] = m1; hm[2][py][ox] = m2; hm[3][py][ox] = m3; hm[4][py][ox] = m4;
    }
    __syncthreads();
    double acc = 0.0;
    for (int i = threadIdx.x; i < SS_T * SS_T; i += 256) {
        const int oy = i / SS_T, ox = i - oy * SS_T;
        const int gy = y0 + oy, gx = x0 + ox;
        if (gy < SS_R || gy >= h - SS_R || gx < SS_R || gx >= w - SS_R) continue;      // the cropped border of the index map
        float m[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int t = 0; t < 11; t++) {
            const float wt = g[t];
#pragma unroll
            for (int q = 0; q < 5; q++) m[q] += wt * hm[q][oy + t][ox];
        }
        const float mu_a = m[0], mu_b = m[1];
        const float va = m[2] - mu_a * mu_a, vb = m[3] - mu_b * mu_b, cab = m[4] - mu_a * mu_b;
        const float idx = ((2.f * mu_a * mu_b + c1) * (2.f * cab + c2)) / ((mu_a * mu_a + mu_b * mu_b + c1) * (va + vb + c2));
        acc += (double)idx;
    }
    acc = warp_sum_d(acc);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int i = 0; i < 8; i++) t += red[i];
        atomicAdd(out, t);
    }
}

// Angle (degrees) between the unit normals F.normalize([gx, gy, nz]) of two touch maps [n][2][h][w] (compute_normal with
// scale_nz, model_utils.py:418-425; torch.cosine_similarity eps 1e-6, clamp to [-1, 1], acos: normal_losses.py:17-33);
// `fake` is clamped to [0, 1] first when clamp_fake (model_utils.py:520).  out[0] += sum of the angles.
__global__ void __launch_bounds__(256) normal_angle_kernel(const float* __restrict__ real, const float* __restrict__ fake, int n, int hw,
                                                           float nz, int clamp_fake, double* out) {
    double acc = 0.0;
    const long long total = (long long)n * hw;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const long long img = i / hw, p = i - img * hw;
        const float rx = real[(img * 2) * hw + p], ry = real[(img * 2 + 1) * hw + p];
        float fx = fake[(img * 2) * hw + p], fy = fake[(img * 2 + 1) * hw + p];
        if (clamp_fake) { fx = fminf(fmaxf(fx, 0.f), 1.f); fy = fminf(fmaxf(fy, 0.f), 1.f); }
        const float rn = fmaxf(sqrtf(rx * rx + ry * ry + nz * nz), 1e-12f), fn = fmaxf(sqrtf(fx * fx + fy * fy + nz * nz), 1e-12f);
        const float r0 = rx / rn, r1 = ry / rn, r2 = nz / rn, f0 = fx / fn, f1 = fy / fn, f2 = nz / fn;
        const float na = fmaxf(sqrtf(r0 * r0 + r1 * r1 + r2 * r2), 1e-6f), nb = fmaxf(sqrtf(f0 * f0 + f1 * f1 + f2 * f2), 1e-6f);
        float c = (r0 * f0 + r1 * f1 + r2 * f2) / (na * nb);
        c = fminf(fmaxf(c, -1.f), 1.f);
        acc += (double)(acosf(c) * 57.29577951308232f);
    }
    acc = warp_sum_d(acc);
    if ((threadIdx.x & 31) == 0) atomicAdd(out, acc);
}

static inline int blocks_for(long long n) {
    long long b = (n + 255) / 256;
    return (int)(b < 1 ? 1 : (b > 148 * 8 ? 148 * 8 : b));
}

}  // namespace skit

using namespace skit;

extern "C" int skit_metric_minmax(const float* x, long long n, float* out2, void* stream) {
    SKIT_REQUIRE(x && out2 && n > 0, "metric_minmax: bad arguments");
    minmax_kernel<<<blocks_for(n), 256, 0, as_stream(stream)>>>(x, n, out2);
    return check_launch("minmax_kernel");
}

extern "C" int skit_metric_sq_err(const float* a, const float* b, long long n, const float* minmax, int clamp_b, double* out, void* stream) {
    SKIT_REQUIRE(a && b && out && n > 0, "metric_sq_err: bad arguments");
    sq_err_kernel<<<blocks_for(n), 256, 0, as_stream(stream)>>>(a, b, n, minmax, clamp_b, out);
    return check_launch("sq_err_kernel");
}

extern "C" int skit_metric_ssim(const float* a, const float* b, int planes, int h, int w, const float* minmax, float data_range,
                                double* out, void* stream) {
    SKIT_REQUIRE(a && b && out && planes > 0 && h > 10 && w > 10, "metric_ssim: images must be larger than the 11x11 window");
    const float c1 = (0.01f * data_range) * (0.01f * data_range), c2 = (0.03f * data_range) * (0.03f * data_range);
    dim3 grid(cdiv(w, SS_T), cdiv(h, SS_T), planes);
    ssim_kernel<<<grid, 256, 0, as_stream(stream)>>>(a, b, h, w, minmax, c1, c2, out);
    return check_launch("ssim_kernel");
}

extern "C" int skit_metric_normal_angle(const float* real, const float* fake, int n, int h, int w, float scale_nz, int clamp_fake,
                                        double* out, void* stream) {
    SKIT_REQUIRE(real && fake && out && n > 0 && h > 0 && w > 0, "metric_normal_angle: bad arguments");
    normal_angle_kernel<<<blocks_for((long long)n * h * w), 256, 0, as_stream(stream)>>>(real, fake, n, h * w, scale_nz, clamp_fake, out);
    return check_launch("normal_angle_kernel");
}
