// Data-pipeline kernels (data/singleskit_dataset.py, data/dataset_util.py): the byte / integer work the
// reference does on the host with Pillow, NumPy and OpenCV for every augmentation of its one (sketch, image, mask) triple:
//   * 8-bit image resize, bit-identical with Pillow's two-pass fixed-point resampler (Image.resize(.., LANCZOS) in zoom_img /
//     crop_img / make_power_2_img, dataset_util.py:159-231)
//   * crop + ToTensor + Normalize(0.5, 0.5) of the cached uint8 sources into fp32 CHW tensors (singleskit_dataset.py:301-315)
//   * the contact-centre search of process_all_valid_patches (singleskit_dataset.py:768-803): which pixels of a touch patch's
//     centre mask have a 32 x 32 window where contact mask x object mask reaches 1 — a Python loop over every centre pixel
//     with a PIL crop inside (the bulk of the reference's 20-30 min start-up); here a hit map, two window-any passes and an
//     ordered compaction, batched over every touch patch of the material
//   * the 32 x 32 gathers of (gx, gy) and of the contact mask for the sampled centres
//   * variance of the Laplacian of the sketch patches (util/util.py:261-265) = the resampling weights
// All HBM / latency bound, integer or fp64 exact: results are compared bit for bit with Pillow / NumPy / OpenCV.
#include <cmath>
#include <vector>
#include "skit_common.cuh"

namespace skit {

constexpr int RS_PRECISION_BITS = 32 - 8 - 2;

// ------------------------------------------------------------------------------------------------- resize coefficients (host)
// Resample.c precompute_coeffs + normalize_coeffs_8bpc for the full source box, in double on the host exactly as Pillow does.
static double rs_sinc(double x) {
    if (x == 0.0) return 1.0;
    x = x * M_PI;
    return sin(x) / x;
}
static double rs_filter(int kind, double x) {
    switch (kind) {
        case SKIT_RESAMPLE_LANCZOS: return (-3.0 <= x && x < 3.0) ? rs_sinc(x) * rs_sinc(x / 3) : 0.0;
        case SKIT_RESAMPLE_BICUBIC: {
            const double a = -0.5;
            if (x < 0.0) x = -x;
            if (x < 1.0) return ((a + 2.0) * x - (a + 3.0)) * x * x + 1;
            if (x < 2.0) return (((x - 5) * x + 8) * x - 4) * a;
            return 0.0;
        }
        case SKIT_RESAMPLE_BILINEAR: if (x < 0.0) x = -x; return x < 1.0 ? 1.0 - x : 0.0;
        case SKIT_RESAMPLE_BOX: return (x > -0.5 && x <= 0.5) ? 1.0 : 0.0;
        case SKIT_RESAMPLE_HAMMING:
            if (x < 0.0) x = -x;
            if (x == 0.0) return 1.0;
            if (x >= 1.0) return 0.0;
            x = x * M_PI;
            return sin(x) / x * (0.54 + 0.46 * cos(x));
    }
    return 0.0;
}
static double rs_support(int kind) {
    switch (kind) {
        case SKIT_RESAMPLE_LANCZOS: return 3.0;
        case SKIT_RESAMPLE_BICUBIC: return 2.0;
        case SKIT_RESAMPLE_BILINEAR: return 1.0;
        case SKIT_RESAMPLE_BOX: return 0.5;
        case SKIT_RESAMPLE_HAMMING: return 1.0;
    }
    return -1.0;
}
// table layout: [out][2 + ksize] ints = {xmin, xmax, k[0..ksize)}
static int rs_coeffs(int in_size, int out_size, int kind, std::vector<int>& table) {
    double scale, filterscale;
    scale = filterscale = (double)in_size / out_size;
    if (filterscale < 1.0) filterscale = 1.0;
    const double support = rs_support(kind) * filterscale;
    const int ksize = (int)ceil(support) * 2 + 1;
    table.assign((size_t)out_size * (2 + ksize), 0);
    std::vector<double> k(ksize);
    const double ss = 1.0 / filterscale;
    for (int xx = 0; xx < out_size; xx++) {
        const double center = (xx + 0.5) * scale;
        double ww = 0.0;
        int xmin = (int)(center - support + 0.5);
        if (xmin < 0) xmin = 0;
        int xmax = (int)(center + support + 0.5);
        if (xmax > in_size) xmax = in_size;
        xmax -= xmin;
        for (int x = 0; x < xmax; x++) {
            const double w = rs_filter(kind, (x + xmin - center + 0.5) * ss);
            k[x] = w;
            ww += w;
        }
        int* row = table.data() + (size_t)xx * (2 + ksize);
        row[0] = xmin; row[1] = xmax;
        for (int x = 0; x < xmax; x++) {
            const double v = ww != 0.0 ? k[x] / ww : k[x];
            row[2 + x] = v < 0 ? (int)(-0.5 + v * (1 << RS_PRECISION_BITS)) : (int)(0.5 + v * (1 << RS_PRECISION_BITS));
        }
    }
    return ksize;
}

__device__ __forceinline__ unsigned char rs_clip8(int v) {
    v >>= RS_PRECISION_BITS;
    return (unsigned char)(v < 0 ? 0 : (v > 255 ? 255 : v));
}

// horizontal pass: one thread per output pixel, all bands at once (the bands of a pixel share bounds and coefficients); consecutive
// threads take consecutive output pixels of a row, so a warp reads one contiguous source segment
template <int C>
__global__ void __launch_bounds__(256) resize_h_kernel(const unsigned char* __restrict__ src, int rows, int sw, unsigned char* __restrict__ dst,
                                                       int dw, const int* __restrict__ tab, int ksize) {
    const int total = rows * dw;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int y = i / dw, xx = i - y * dw;
        const int* row = tab + xx * (2 + ksize);
        const int xmin = row[0], xmax = row[1];
        int acc[C];
#pragma unroll
        for (int ch = 0; ch < C; ch++) acc[ch] = 1 << (RS_PRECISION_BITS - 1);
        const unsigned char* s = src + ((size_t)y * sw + xmin) * C;
        for (int x = 0; x < xmax; x++) {
            const int k = row[2 + x];
#pragma unroll
            for (int ch = 0; ch < C; ch++) acc[ch] += (int)s[x * C + ch] * k;
        }
#pragma unroll
        for (int ch = 0; ch < C; ch++) dst[(size_t)i * C + ch] = rs_clip8(acc[ch]);
    }
}
// vertical pass: a thread owns 4 consecutive bytes of an output row (one 32-bit load per tap, one 32-bit store); rows are wc bytes,
// wc4 = wc / 4 words (the caller pads nothing: wc % 4 != 0 takes the byte kernel below)
__global__ void __launch_bounds__(256) resize_v4_kernel(const unsigned int* __restrict__ src, int wc4, unsigned int* __restrict__ dst, int dh,
                                                        const int* __restrict__ tab, int ksize) {
    const int total = dh * wc4;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < total; i += gridDim.x * blockDim.x) {
        const int yy = i / wc4, xq = i - yy * wc4;
        const int* row = tab + yy * (2 + ksize);
        const int ymin = row[0], ymax = row[1];
        int a0 = 1 << (RS_PRECISION_BITS - 1), a1 = a0, a2 = a0, a3 = a0;
        for (int y = 0; y < ymax; y++) {
            const unsigned int v = src[(size_t)(ymin + y) * wc4 + xq];
            const int k = row[2 + y];
            a0 += (int)(v & 255u) * k; a1 += (int)((v >> 8) & 255u) * k; a2 += (int)((v >> 16) & 255u) * k; a3 += (int)(v >> 24) * k;
        }
        dst[i] = (unsigned int)rs_clip8(a0) | ((unsigned int)rs_clip8(a1) << 8) | ((unsigned int)rs_clip8(a2) << 16) | ((unsigned int)rs_clip8(a3) << 24);
    }
}
__global__ void __launch_bounds__(256) resize_v_kernel(const unsigned char* __restrict__ src, int sh, int wc,
                                                       unsigned char* __restrict__ dst, int dh, const int* __restrict__ tab, int ksize) {
    const long long total = (long long)dh * wc;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int yy = (int)(i / wc);
        const int xc = (int)(i % wc);
        const int* row = tab + (long long)yy * (2 + ksize);
        const int ymin = row[0], ymax = row[1];
        int acc = 1 << (RS_PRECISION_BITS - 1);
        for (int y = 0; y < ymax; y++) acc += (int)src[(long long)(ymin + y) * wc + xc] * row[2 + y];
        dst[i] = rs_clip8(acc);
    }
}

static int grid_for(long long n) { return (int)std::min<long long>(cdivll(n, 256), 148LL * 16); }

// ------------------------------------------------------------------------------------------------- crop -> fp32 CHW
// a thread produces 4 consecutive pixels of one output row for every band: C x 4 source bytes in (contiguous), one 16-byte store per
// band out.  grid: (ceil(w / 4 / 256), h)
template <int C>
__global__ void __launch_bounds__(256) crop_to_tensor_kernel(const unsigned char* __restrict__ src, int sh, int sw, int y0, int x0,
                                                             int h, int w, int normalize, float* __restrict__ dst) {
    const int y = blockIdx.y, xq = blockIdx.x * blockDim.x + threadIdx.x;
    const int x = xq * 4;
    if (x >= w) return;
    const int sy = y + y0;
    float v[C][4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
        const int sx = x + k + x0;
        const bool in = sy >= 0 && sy < sh && sx >= 0 && sx < sw && x + k < w;            // PIL crop pads with 0
        const unsigned char* s = src + ((size_t)sy * sw + sx) * C;
#pragma unroll
        for (int ch = 0; ch < C; ch++) {
            float f = __fdiv_rn(in ? (float)s[ch] : 0.f, 255.f);                          // ToTensor: byte -> float, .div(255)
            if (normalize) f = __fdiv_rn(__fsub_rn(f, 0.5f), 0.5f);                       // Normalize(0.5, 0.5)
            v[ch][k] = f;
        }
    }
#pragma unroll
    for (int ch = 0; ch < C; ch++) {
        float* d = dst + ((size_t)ch * h + y) * w + x;
        if ((w & 3) == 0) *reinterpret_cast<float4*>(d) = make_float4(v[ch][0], v[ch][1], v[ch][2], v[ch][3]);
        else
            for (int k = 0; k < 4 && x + k < w; k++) d[k] = v[ch][k];
    }
}

// ------------------------------------------------------------------------------------------------- contact centres
struct TouchSet {
    const double* touch_mask;          // all patches back to back, row-major, values as the reference holds them after its /255
    const unsigned char* center_mask;  // > 0 where the reference's touch_center_mask > 0
    const long long* pix_off;          // [P + 1]
    const int* ph; const int* pw;      // [P]
    const int* roi_x; const int* roi_y;  // [P] position of each patch in the augmented image (new_ROI_x3 / y3)
    int P;
    const unsigned char* M; int mh, mw;  // object mask of the augmentation (M3), single channel
};

__device__ __forceinline__ int find_patch(const long long* off, int P, long long i) {
    int lo = 0, hi = P - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (off[mid] <= i) lo = mid; else hi = mid - 1;
    }
    return lo;
}
__device__ __forceinline__ double mask_at(const TouchSet& t, int p, int y, int x) {
    const int my = t.roi_y[p] + y, mx = t.roi_x[p] + x;
    return (my >= 0 && my < t.mh && mx >= 0 && mx < t.mw) ? (double)t.M[(long long)my * t.mw + mx] : 0.0;
}
// hit[i] = (touch_mask x M / 255 >= 1) at pixel i, in the reference's own fp64 arithmetic
__global__ void __launch_bounds__(256) contact_hit_kernel(TouchSet t, unsigned char* __restrict__ hit) {
    const long long total = t.pix_off[t.P];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int p = find_patch(t.pix_off, t.P, i);
        const int r = (int)(i - t.pix_off[p]);
        const int y = r / t.pw[p], x = r % t.pw[p];
        const double v = __ddiv_rn(__dmul_rn(t.touch_mask[i], mask_at(t, p, y, x)), 255.0);
        hit[i] = v >= 1.0;
    }
}
// rowany[i] = any hit in [x - half, x + half) of the same row
__global__ void __launch_bounds__(256) contact_rowany_kernel(TouchSet t, const unsigned char* __restrict__ hit, int half,
                                                             unsigned char* __restrict__ rowany) {
    const long long total = t.pix_off[t.P];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int p = find_patch(t.pix_off, t.P, i);
        const int w = t.pw[p];
        const int x = (int)((i - t.pix_off[p]) % w);
        unsigned char any = 0;
        const int lo = max(0, x - half), hi = min(w, x + half);
        for (int q = lo; q < hi; q++) any |= hit[i + (q - x)];
        rowany[i] = any;
    }
}
// flag[i] = centre mask set and any rowany in rows [y - half, y + half)
__global__ void __launch_bounds__(256) contact_flag_kernel(TouchSet t, const unsigned char* __restrict__ rowany, int half,
                                                           unsigned char* __restrict__ flag) {
    const long long total = t.pix_off[t.P];
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        unsigned char f = 0;
        if (t.center_mask[i]) {
            const int p = find_patch(t.pix_off, t.P, i);
            const int w = t.pw[p], h = t.ph[p];
            const int y = (int)((i - t.pix_off[p]) / w);
            const int lo = max(0, y - half), hi = min(h, y + half);
            for (int q = lo; q < hi && !f; q++) f |= rowany[i + (long long)(q - y) * w];
        }
        flag[i] = f;
    }
}
// one block per patch: ordered (row-major, as np.where lists them) compaction of the flagged pixels; also whether the patch's
// rectangle touches the object mask at all (singleskit_dataset.py:745)
__global__ void __launch_bounds__(256) contact_compact_kernel(TouchSet t, const unsigned char* __restrict__ flag, int* __restrict__ counts,
                                                              int* __restrict__ centers, int* __restrict__ in_mask) {
    __shared__ int warp_tot[8];
    __shared__ int base_s, any_s;
    const int p = blockIdx.x;
    const long long off = t.pix_off[p];
    const int n = t.ph[p] * t.pw[p];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    if (threadIdx.x == 0) { base_s = 0; any_s = 0; }
    __syncthreads();
    int any = 0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const int y = i / t.pw[p], x = i % t.pw[p];
        if (mask_at(t, p, y, x) != 0.0) any = 1;
    }
    if (any) atomicOr(&any_s, 1);
    for (int i0 = 0; i0 < n; i0 += 256) {
        const int i = i0 + threadIdx.x;
        const int f = (i < n) ? flag[off + i] : 0;
        const unsigned bal = __ballot_sync(0xffffffffu, f);
        if (lane == 0) warp_tot[warp] = __popc(bal);
        __syncthreads();
        int before = base_s;
        for (int q = 0; q < warp; q++) before += warp_tot[q];
        if (f) centers[off + before + __popc(bal & ((1u << lane) - 1))] = i;
        __syncthreads();
        if (threadIdx.x == 0) { int s = 0; for (int q = 0; q < 8; q++) s += warp_tot[q]; base_s += s; }
        __syncthreads();
    }
    if (threadIdx.x == 0) { counts[p] = base_s; in_mask[p] = any_s; }
}

// ------------------------------------------------------------------------------------------------- sampled squares
// one block per selection k = (patch, cx, cy): copies the patch x patch windows of the two gradient maps (raw bytes of `esize`
// each: the npz's own dtype survives) and writes square_mask = touch_mask x M_patch / 255 in fp64
__global__ void __launch_bounds__(256) touch_squares_kernel(TouchSet t, const unsigned char* __restrict__ gx, const unsigned char* __restrict__ gy,
                                                            int esize, const int* __restrict__ sel_patch, const int* __restrict__ sel_cx,
                                                            const int* __restrict__ sel_cy, int patch, unsigned char* __restrict__ t_images,
                                                            double* __restrict__ i_masks) {
    const int k = blockIdx.x;
    const int p = sel_patch[k];
    const int x0 = sel_cx[k] - patch / 2, y0 = sel_cy[k] - patch / 2;
    const int w = t.pw[p];
    const long long off = t.pix_off[p];
    const int n = patch * patch;
    for (int i = threadIdx.x; i < n; i += 256) {
        const int y = i / patch, x = i % patch;
        const long long s = off + (long long)(y0 + y) * w + (x0 + x);
        for (int b = 0; b < esize; b++) {
            t_images[(((long long)k * 2 + 0) * n + i) * esize + b] = gx[s * esize + b];
            t_images[(((long long)k * 2 + 1) * n + i) * esize + b] = gy[s * esize + b];
        }
        i_masks[(long long)k * n + i] = __ddiv_rn(__dmul_rn(t.touch_mask[s], mask_at(t, p, y0 + y, x0 + x)), 255.0);
    }
}

// ------------------------------------------------------------------------------------------------- Laplacian variance
// one block per window: (image - ref) wraps in uint8 as NumPy does, 4-neighbour Laplacian with reflect-101 borders, population
// variance.  Everything is an integer up to the last division, so the sums are exact: var = (n * S2 - S1^2) / n^2.
__global__ void __launch_bounds__(256) laplacian_var_kernel(const unsigned char* __restrict__ img, int h, int w, const int* __restrict__ x0s,
                                                            const int* __restrict__ y0s, int size, int ref, double* __restrict__ out) {
    extern __shared__ int win[];     // size x size wrapped values
    __shared__ long long red[2][8];
    const int k = blockIdx.x;
    const int x0 = x0s[k], y0 = y0s[k];
    const int n = size * size;
    for (int i = threadIdx.x; i < n; i += 256) {
        const int y = y0 + i / size, x = x0 + i % size;
        const int v = (y >= 0 && y < h && x >= 0 && x < w) ? img[(long long)y * w + x] : 0;       // PIL crop pads with 0
        win[i] = (v - ref) & 255;
    }
    __syncthreads();
    long long s1 = 0, s2 = 0;
    for (int i = threadIdx.x; i < n; i += 256) {
        const int y = i / size, x = i % size;
        const int ym = y == 0 ? (size > 1 ? 1 : 0) : y - 1, yp = y == size - 1 ? (size > 1 ? size - 2 : 0) : y + 1;
        const int xm = x == 0 ? (size > 1 ? 1 : 0) : x - 1, xp = x == size - 1 ? (size > 1 ? size - 2 : 0) : x + 1;
        const int lap = win[ym * size + x] + win[yp * size + x] + win[y * size + xm] + win[y * size + xp] - 4 * win[i];
        s1 += lap; s2 += (long long)lap * lap;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) { s1 += __shfl_xor_sync(0xffffffffu, s1, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if ((threadIdx.x & 31) == 0) { red[0][threadIdx.x >> 5] = s1; red[1][threadIdx.x >> 5] = s2; }
    __syncthreads();
    if (threadIdx.x == 0) {
        long long a = 0, b = 0;
        for (int q = 0; q < 8; q++) { a += red[0][q]; b += red[1][q]; }
        // np.var: mean of (x - mean)^2 in fp64; the exact rational value, rounded once
        const double nn = (double)n;
        out[k] = ((double)b - (double)a * (double)a / nn) / nn;
    }
}

// ------------------------------------------------------------------------------------------------- random-patch candidates
// get_patch_in_input's random mode (models/model_utils.py:212-218): clamp(conv2d(M, ones(k, k), padding=pad), 0, 1) != 0, i.e. for a
// non-negative mask "some mask pixel in the k x k window", on the (h + 2 pad - k + 1) x (w + 2 pad - k + 1) map whose nonzero
// positions are the candidate offsets.  Separable window-any; the result leaves as a bit map (bit c % 32 of word c / 32, row-major)
// plus per-row counts, so the host keeps (h - 14) x (w - 14) / 8 bytes instead of pulling the mask over and running cumulative sums.
__global__ void __launch_bounds__(256) mask_rowwin_kernel(const float* __restrict__ M, int h, int w, int k, int pad, int ow,
                                                          unsigned char* __restrict__ rowwin) {
    const long long total = (long long)h * ow;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total; i += (long long)gridDim.x * blockDim.x) {
        const int y = (int)(i / ow), c = (int)(i % ow);
        const int lo = max(0, c - pad), hi = min(w, c - pad + k);
        unsigned char any = 0;
        for (int x = lo; x < hi; x++) any |= (M[(long long)y * w + x] > 0.f);
        rowwin[i] = any;
    }
}
// one block per output row r: box[r][c] = any(rowwin[r - pad .. r - pad + k - 1][c])
__global__ void __launch_bounds__(256) mask_box_bits_kernel(const unsigned char* __restrict__ rowwin, int h, int k, int pad, int ow, int words,
                                                            unsigned int* __restrict__ bits, int* __restrict__ rowcount) {
    __shared__ int cnt;
    const int r = blockIdx.x;
    if (threadIdx.x == 0) cnt = 0;
    __syncthreads();
    const int lo = max(0, r - pad), hi = min(h, r - pad + k);
    int mine = 0;
    for (int c0 = 0; c0 < words * 32; c0 += 256) {
        const int c = c0 + threadIdx.x;
        int any = 0;
        if (c < ow)
            for (int y = lo; y < hi && !any; y++) any = rowwin[(long long)y * ow + c];
        const unsigned bal = __ballot_sync(0xffffffffu, any);
        if ((threadIdx.x & 31) == 0 && (c >> 5) < words) { bits[(long long)r * words + (c >> 5)] = bal; mine += __popc(bal); }
    }
    if (mine) atomicAdd(&cnt, mine);
    __syncthreads();
    if (threadIdx.x == 0) rowcount[r] = cnt;
}

}  // namespace skit

using namespace skit;

extern "C" int skit_mask_box_bits(const float* M, int h, int w, int k, int pad, unsigned char* scratch, unsigned int* bits, int* rowcount,
                                  void* stream) {
    const int oh = h + 2 * pad - k + 1, ow = w + 2 * pad - k + 1;
    SKIT_REQUIRE(M && scratch && bits && rowcount && k >= 1 && pad >= 0 && oh > 0 && ow > 0, "mask_box_bits: bad arguments (%d x %d, k %d, pad %d)", h, w, k, pad);
    const int words = (ow + 31) / 32;
    mask_rowwin_kernel<<<grid_for((long long)h * ow), 256, 0, as_stream(stream)>>>(M, h, w, k, pad, ow, scratch);
    mask_box_bits_kernel<<<oh, 256, 0, as_stream(stream)>>>(scratch, h, k, pad, ow, words, bits, rowcount);
    return check_launch("mask_box_bits");
}

namespace {
// device copies of the coefficient tables, keyed by (in, out, filter): a dataset resizes a handful of shapes again and again, and the
// table build (sin() in double per tap) plus its upload cost more than the kernels
struct RsTable { int in, out, kind, ksize; int* dev; };
constexpr int RS_CACHE = 16;
RsTable g_rs_tables[RS_CACHE];       // zero-initialised: dev == nullptr marks a free slot
int g_rs_next = 0;
int rs_table(int in, int out, int kind, cudaStream_t st, const int** dev) {
    for (const RsTable& e : g_rs_tables)
        if (e.dev && e.in == in && e.out == out && e.kind == kind) { *dev = e.dev; return e.ksize; }
    std::vector<int> t;
    const int ksize = rs_coeffs(in, out, kind, t);
    int* d = nullptr;
    if (cudaMalloc(&d, t.size() * sizeof(int)) != cudaSuccess) return -1;
    cudaMemcpyAsync(d, t.data(), t.size() * sizeof(int), cudaMemcpyHostToDevice, st);     // pageable source: staged before the call returns
    RsTable& e = g_rs_tables[g_rs_next++ % RS_CACHE];                                     // round-robin replacement
    if (e.dev) { cudaStreamSynchronize(st); cudaFree(e.dev); }                            // a kernel in flight may still read the old table
    e = RsTable{in, out, kind, ksize, d};
    *dev = d;
    return ksize;
}
}  // namespace

extern "C" int skit_resize_u8(const unsigned char* src, int sh, int sw, int c, unsigned char* dst, int dh, int dw, int filter, void* stream) {
    SKIT_REQUIRE(src && dst && sh > 0 && sw > 0 && dh > 0 && dw > 0 && c >= 1 && c <= 4, "resize_u8: bad shape %dx%dx%d -> %dx%d", sh, sw, c, dh, dw);
    SKIT_REQUIRE(rs_support(filter) > 0, "resize_u8: unsupported filter %d (BOX 4, BILINEAR 2, HAMMING 5, BICUBIC 3, LANCZOS 1)", filter);
    SKIT_REQUIRE((long long)sh * dw < (1ll << 31) / 4 && (long long)dh * dw * c < (1ll << 31), "resize_u8: image too large for 32-bit indexing");
    cudaStream_t st = as_stream(stream);
    if (sh == dh && sw == dw) {      // Image.resize returns a copy
        cudaMemcpyAsync(dst, src, (size_t)sh * sw * c, cudaMemcpyDeviceToDevice, st);
        return check_launch("resize_u8 copy");
    }
    const int *d_th = nullptr, *d_tv = nullptr;
    int kh = 0, kv = 0;
    if (sw != dw) { kh = rs_table(sw, dw, filter, st, &d_th); SKIT_REQUIRE(kh > 0, "resize_u8: coefficient table allocation failed"); }
    if (sh != dh) { kv = rs_table(sh, dh, filter, st, &d_tv); SKIT_REQUIRE(kv > 0, "resize_u8: coefficient table allocation failed"); }
    // Pillow's horizontal pass only covers the source rows the vertical pass reads; the result is the same as covering all rows
    unsigned char* tmp = nullptr;
    const unsigned char* vsrc = src;
    if (kh) {
        unsigned char* hout = dst;
        if (kv) { cudaMallocAsync(&tmp, (size_t)sh * dw * c, st); hout = tmp; }
        const int g = grid_for((long long)sh * dw);
        switch (c) {
            case 1: resize_h_kernel<1><<<g, 256, 0, st>>>(src, sh, sw, hout, dw, d_th, kh); break;
            case 2: resize_h_kernel<2><<<g, 256, 0, st>>>(src, sh, sw, hout, dw, d_th, kh); break;
            case 3: resize_h_kernel<3><<<g, 256, 0, st>>>(src, sh, sw, hout, dw, d_th, kh); break;
            default: resize_h_kernel<4><<<g, 256, 0, st>>>(src, sh, sw, hout, dw, d_th, kh); break;
        }
        vsrc = hout;
    }
    if (kv) {
        const int wc = dw * c;
        // 32-bit path: rows must start on 4-byte boundaries in both buffers (cudaMalloc'd bases are 256-byte aligned; row pitch wc)
        if (wc % 4 == 0 && ((uintptr_t)vsrc & 3) == 0 && ((uintptr_t)dst & 3) == 0)
            resize_v4_kernel<<<grid_for((long long)dh * (wc / 4)), 256, 0, st>>>((const unsigned int*)vsrc, wc / 4, (unsigned int*)dst, dh, d_tv, kv);
        else
            resize_v_kernel<<<grid_for((long long)dh * wc), 256, 0, st>>>(vsrc, sh, wc, dst, dh, d_tv, kv);
    }
    const int rc = check_launch("resize_u8");
    if (tmp) cudaFreeAsync(tmp, st);
    return rc;
}

extern "C" int skit_u8_crop_to_tensor(const unsigned char* src, int sh, int sw, int c, int y0, int x0, int h, int w, int normalize,
                                      float* dst, void* stream) {
    SKIT_REQUIRE(src && dst && sh > 0 && sw > 0 && h > 0 && w > 0 && c >= 1 && c <= 4, "u8_crop_to_tensor: bad shape");
    dim3 grid(cdiv(cdiv(w, 4), 256), h);
    cudaStream_t st = as_stream(stream);
    switch (c) {
        case 1: crop_to_tensor_kernel<1><<<grid, 256, 0, st>>>(src, sh, sw, y0, x0, h, w, normalize, dst); break;
        case 2: crop_to_tensor_kernel<2><<<grid, 256, 0, st>>>(src, sh, sw, y0, x0, h, w, normalize, dst); break;
        case 3: crop_to_tensor_kernel<3><<<grid, 256, 0, st>>>(src, sh, sw, y0, x0, h, w, normalize, dst); break;
        default: crop_to_tensor_kernel<4><<<grid, 256, 0, st>>>(src, sh, sw, y0, x0, h, w, normalize, dst); break;
    }
    return check_launch("crop_to_tensor_kernel");
}

static TouchSet make_set(const double* touch_mask, const unsigned char* center_mask, const long long* pix_off, const int* ph, const int* pw,
                         const int* roi_x, const int* roi_y, int P, const unsigned char* M, int mh, int mw) {
    TouchSet t;
    t.touch_mask = touch_mask; t.center_mask = center_mask; t.pix_off = pix_off; t.ph = ph; t.pw = pw;
    t.roi_x = roi_x; t.roi_y = roi_y; t.P = P; t.M = M; t.mh = mh; t.mw = mw;
    return t;
}

extern "C" int skit_contact_centers(const double* touch_mask, const unsigned char* center_mask, const long long* pix_off, long long total_pixels,
                                    const int* ph, const int* pw, const int* roi_x, const int* roi_y, int P, const unsigned char* M, int mh,
                                    int mw, int patch, unsigned char* scratch, int* counts, int* centers, int* in_mask, void* stream) {
    SKIT_REQUIRE(touch_mask && center_mask && pix_off && ph && pw && roi_x && roi_y && M && scratch && counts && centers && in_mask,
                 "contact_centers: null argument");
    SKIT_REQUIRE(P > 0 && total_pixels > 0 && patch >= 2 && patch % 2 == 0, "contact_centers: bad sizes (P %d, patch %d)", P, patch);
    const TouchSet t = make_set(touch_mask, center_mask, pix_off, ph, pw, roi_x, roi_y, P, M, mh, mw);
    cudaStream_t st = as_stream(stream);
    unsigned char *hit = scratch, *rowany = scratch + total_pixels, *flag = scratch + 2 * total_pixels;
    const int g = grid_for(total_pixels);
    contact_hit_kernel<<<g, 256, 0, st>>>(t, hit);
    contact_rowany_kernel<<<g, 256, 0, st>>>(t, hit, patch / 2, rowany);
    contact_flag_kernel<<<g, 256, 0, st>>>(t, rowany, patch / 2, flag);
    contact_compact_kernel<<<P, 256, 0, st>>>(t, flag, counts, centers, in_mask);
    return check_launch("contact_centers");
}

extern "C" int skit_touch_squares(const double* touch_mask, const long long* pix_off, const int* ph, const int* pw, const int* roi_x,
                                  const int* roi_y, int P, const unsigned char* M, int mh, int mw, const void* gx, const void* gy, int elem_size,
                                  const int* sel_patch, const int* sel_cx, const int* sel_cy, int K, int patch, void* t_images, double* i_masks,
                                  void* stream) {
    SKIT_REQUIRE(touch_mask && pix_off && pw && roi_x && roi_y && M && gx && gy && sel_patch && sel_cx && sel_cy && t_images && i_masks,
                 "touch_squares: null argument");
    SKIT_REQUIRE(K > 0 && (elem_size == 4 || elem_size == 8) && patch >= 2, "touch_squares: bad sizes (K %d, element %d bytes)", K, elem_size);
    const TouchSet t = make_set(touch_mask, nullptr, pix_off, ph, pw, roi_x, roi_y, P, M, mh, mw);
    touch_squares_kernel<<<K, 256, 0, as_stream(stream)>>>(t, (const unsigned char*)gx, (const unsigned char*)gy, elem_size, sel_patch, sel_cx,
                                                           sel_cy, patch, (unsigned char*)t_images, i_masks);
    return check_launch("touch_squares_kernel");
}

extern "C" int skit_laplacian_var_u8(const unsigned char* img, int h, int w, const int* x0, const int* y0, int K, int size, int ref, double* out,
                                     void* stream) {
    SKIT_REQUIRE(img && x0 && y0 && out && K > 0 && size >= 1 && size <= 96, "laplacian_var_u8: bad arguments (K %d, size %d)", K, size);
    laplacian_var_kernel<<<K, 256, (size_t)size * size * sizeof(int), as_stream(stream)>>>(img, h, w, x0, y0, size, ref, out);
    return check_launch("laplacian_var_kernel");
}
