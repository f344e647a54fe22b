/* skit_b200 — C-ABI of the B200-native skitG/sinskitG hot path.
 *
 * The reference (RuihanGao/visual-tactile-synthesis) is pure PyTorch: its "FFI" for this path
 * is ATen (F.conv2d, F.instance_norm, F.batch_norm, advanced indexing, Adam ...).  Every entry
 * point below names the reference call site (file:line under /root/reference) it replaces.
 * All pointers are DEVICE pointers unless stated otherwise; `stream` is a cudaStream_t passed as
 * void*.  Every function returns SKIT_OK (0) or a negative error code; skit_last_error() returns
 * the message of the last failure on the calling thread.  No function synchronises the device.
 *
 * Device data layout:
 *   feature maps ......... NHWC fp32 ("raw" conv outputs, residual stream, gradients)
 *   conv operands ........ NHWC with an explicit halo, either fp32 (SKIT_FMT_F32) or two bf16
 *                          planes hi/lo with hi+lo ~= x (SKIT_FMT_BF16X2) for the tcgen05 path
 *   images / patches ..... NCHW fp32 planar, exactly the reference's tensors
 *   parameters ........... reference layout ([Co][Ci][kh][kw] fp32), repacked per step
 */
#ifndef SKIT_B200_H
#define SKIT_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SKIT_OK 0
#define SKIT_ERR_INVALID (-1)
#define SKIT_ERR_CUDA (-2)
#define SKIT_ERR_UNSUPPORTED (-3)

#define SKIT_FMT_F32 0
#define SKIT_FMT_BF16X2 1

#define SKIT_PAD_ZERO 0
#define SKIT_PAD_REFLECT 1
#define SKIT_PAD_REPLICATE 2

#define SKIT_ACT_NONE 0
#define SKIT_ACT_RELU 1
#define SKIT_ACT_LRELU 2 /* negative slope 0.2 (networks.py:1712, unet_parts_custom.py:23) */

#define SKIT_NORM_NONE 0
#define SKIT_NORM_INSTANCE 1 /* statistics per (n, c) */
#define SKIT_NORM_BATCH 2    /* statistics per c over (n, h, w) */

/* The (sum g, sum g*xhat) buffers of the norm backward hold this many replicas, [R][groups][c][2] doubles: a producer CTA adds
 * into replica (its index mod R), consumers add the replicas up.  Hundreds of CTAs hitting the same few hundred fp64
 * addresses serialise in L2 (measured: 25 of 39 us of the phase-A kernel at the trunk shape); R copies divide that.
 * Full layout in doubles: [R][groups][c][2] partial sums | [groups][c] finals (two floats: sum/count each) | 1 ticket slot —
 * (2R + 1) * groups * c + 1 doubles, zeroed by the caller; the last CTA of phase A writes the finals phase B reads. */
#define SKIT_SUM_REPLICAS 8

#define SKIT_IMPL_AUTO 0
#define SKIT_IMPL_SIMT 1
#define SKIT_IMPL_TC 2 /* tcgen05 + TMA; fails with SKIT_ERR_UNSUPPORTED if the shape is not eligible */

/* A conv operand: NHWC tensor [n][hp][wp][c] that already contains its halo. */
typedef struct skit_operand {
    void* p0;  /* F32: float*;  BF16X2: bf16 hi plane */
    void* p1;  /* BF16X2: bf16 lo plane; otherwise NULL */
    int fmt;   /* SKIT_FMT_* */
    int n, hp, wp, c;
} skit_operand;

/* Packed conv weights produced by skit_pack_conv_weights. */
typedef struct skit_weights {
    const float* f32; /* [k*k*ci][co]  (SIMT path)           */
    const void* hi;   /* bf16 [k*k][co][ci] (tcgen05 path)   */
    const void* lo;   /* bf16 [k*k][co][ci]                  */
    int k;
    int ci; /* channels the GEMM reduces over per tap (forward pack: conv ci; dgrad packs: conv co) */
    int co; /* GEMM output columns               (forward pack: conv co; dgrad packs: conv ci) */
    int kw; /* filter columns; 0 = square (k x k).  1 for x-folded packs (modes 4 / 5): the filter's columns live in the
               operand's channel axis, see skit_fold_x_operand */
} skit_weights;

const char* skit_last_error(void);
/* Library/ABI version and the compute capability the kernels were built for (100 = sm_100a). */
int skit_version(void);
int skit_built_arch(void);

/* ---------------------------------------------------------------- weights
 * Repack a reference-layout conv weight [co][ci][k][k] (nn.Conv2d.weight, networks.py:1078 etc.).
 * mode 0: forward pack        Wf[(ky*k+kx)*ci + c][o]      = w[o][c][ky][kx]
 * mode 1: stride-1 dgrad pack Wd[((k-1-ky)*k+(k-1-kx))*co + o][c] = w[o][c][ky][kx]
 *         (conv of the zero-padded output gradient with the flipped, transposed filter)
 * mode 2: gather dgrad pack   Wg[(ky*k+kx)*co + o][c]      = w[o][c][ky][kx]
 * mode 3: stride-2 phase dgrad pack (k even): tap = phase*(k/2)^2 + (k/2-1-ky/2)*(k/2) + (k/2-1-kx/2),
 *         phase = (ky%2)*2 + kx%2;  Wp[tap*co + o][c] = w[o][c][ky][kx]   (see skit_conv2d_dgrad_s2)
 * f32 / hi / lo may each be NULL to skip that product.  hi/lo are [tap][N][K] K-major bf16.  */
int skit_pack_conv_weights(const float* w, int co, int ci, int k, int mode,
                           float* f32, void* hi, void* lo, void* stream);
/* bf16 hi/lo pack with the reduction axis zero-padded to kpad channels (a multiple of 8 >= the real count): lets layers
 * with thin inputs (the 9-channel generator stem, the 5-channel head's input gradient) run on the tcgen05 path, whose
 * TMA rows must be 16-byte multiples.  The matching operand carries the same number of (zero) padding channels. */
int skit_pack_conv_weights_padded(const float* w, int co, int ci, int k, int mode, int kpad, void* hi, void* lo, void* stream);
/* All packs of a net in ONE launch (after each Adam step): descs_dev is a DEVICE array of n descriptors; a descriptor
 * fills either the fp32 pack (hi == NULL) or the bf16 hi/lo pack (with kpad as in the padded variant).  `start` is the
 * prefix sum of the packs' element counts (k*k*N*K, K padded for bf16), `total` their sum. */
typedef struct skit_pack_desc {
    const float* w;   /* reference-layout weight [co][ci][k][k] */
    float* f32;
    void* hi;
    void* lo;
    long long start;
    int co, ci, k, mode, kpad, reserved;
} skit_pack_desc;
int skit_pack_conv_weights_batched(const skit_pack_desc* descs_dev, int n, long long total, void* stream);
/* Same refresh for packs with k <= 4 (modes 0-3), staged through shared memory in [32 co][32 ci][k*k] bricks so that the
 * reference-layout source is read with unit stride.  Here descs[i].start is the prefix sum of TILES,
 * ceil(N/32) * ceil(Kpadded/32) per pack (N, K as in skit_pack_conv_weights); a pack may carry fp32 and bf16 outputs. */
int skit_pack_conv_weights_tiled(const skit_pack_desc* descs_dev, int n, int total_tiles, int max_k, void* stream);
/* x-folded packs for thin k x k layers on the tensor cores (the 9-channel 7x7 generator stem, the input gradient of the
 * 5-channel 7x7 head): the filter's kw columns move into the 64-wide channel axis, leaving a (k x 1) filter:
 *   mode 4 (forward):        Wf[ky][o][kx*cp + c]  = w[o][c][ky][kx]            (N = co, K = 64, zero beyond kw*cp and c >= ci)
 *   mode 5 (input gradient): Wg[ky][c][kx*cp + o]  = w[o][c][k-1-ky][k-1-kx]    (N = ci, K = 64)
 * cp = channels per folded column (>= ci resp. co, k*cp <= 64).  hi/lo: bf16 [k][N][64]. */
int skit_pack_conv_weights_folded(const float* w, int co, int ci, int k, int mode, int cp, void* hi, void* lo, void* stream);
/* thin: haloed operand [n][hp][wp][cp] (fp32 or bf16x2).  folded: bf16x2 operand [n][hp][wp-kw+1][64] with
 * folded[n][y][x][kx*cp + c] = thin[n][y][x + kx][c] (zero for channels >= kw*cp): a (k x kw) valid conv over `thin` equals a
 * (k x 1) valid conv over `folded` with a mode-4 / mode-5 pack — 64-channel K steps instead of kw thin ones. */
int skit_fold_x_operand(const skit_operand* thin, int kw, const skit_operand* folded, void* stream);
/* Weight gradient against an x-folded input operand: dw[o][c][ky][kx] += sum dy[pix][o] * folded[pix + ky][kx*cp + c].
 * xf: folded operand (c = 64), dy: bf16x2 operand with 64-multiple channels; scratch: k*64*dy->c floats, zeroed. */
int skit_conv2d_wgrad_folded(const skit_operand* xf, int org, const skit_operand* dy, int dy_org, int k, int kw, int cp,
                             int ho, int wo, float* scratch, float* dw, int co_real, int ci_real, void* stream);
/* Weight gradient of a thin-OUTPUT k x k layer against the x-folded GRADIENT operand (the one its input gradient uses):
 *   dw[o][c][ky][kx] += sum_{y, X} dyf[y + k-1][X][(k-1-kx)*cp + o] * x[y + ky][X][c],  y < ho, X < wo + k - 1
 * x: the layer's haloed bf16x2 input operand (64-multiple channels); dyf = skit_fold_x_operand(dy haloed by k-1, k);
 * scratch: k*64*x->c floats, zeroed.  k row taps instead of k*k passes over x. */
int skit_conv2d_wgrad_dyfolded(const skit_operand* x, const skit_operand* dyf, int k, int cp, int ho, int wo,
                               float* scratch, float* dw, int co_real, int ci_real, void* stream);
/* skit_dbias over the first nch channels only (channel-padded gradient operands). */
int skit_dbias_n(const skit_operand* dy, int dy_org, int ho, int wo, int nch, float* dbias, void* stream);
/* Inverse of the forward pack for gradients: dWf [(tap*ci+c)][o] fp32 -> dw[o][c][ky][kx] (+= if accumulate). */
int skit_unpack_conv_wgrad(const float* dwf, int co, int ci, int k, float* dw, int accumulate, void* stream);

/* ---------------------------------------------------------------- convolution
 * Replaces F.conv2d / nn.Conv2d.forward (networks.py:1078,1090,1281-1310,1124,1706-1727) as a
 * VALID convolution over an operand that already carries its halo:
 *   y[n][oy][ox][o] = bias[o] + sum_{ky,kx,c} x[n][org+oy*stride+ky][org+ox*stride+kx][c] * W[ky][kx][c][o]
 * y is NHWC fp32 [n][ho][wo][co].  If stats != NULL, accumulates sum and sum-of-squares of y
 * (double, [groups][co][2]; groups = n for SKIT_NORM_INSTANCE, 1 for SKIT_NORM_BATCH) — the
 * InstanceNorm/BatchNorm statistics, fused into the conv epilogue.  stats must be zeroed by the caller. */
int skit_conv2d_fwd(const skit_operand* x, const skit_weights* w, int stride, int org,
                    int ho, int wo, const float* bias, float* y,
                    double* stats, int stats_mode, int impl, void* stream);

/* Input gradient of the conv above for any stride (autograd of F.conv2d), gather form:
 *   dx[n][iy][ix][c] = sum_{ky,kx,o : (iy-ky)%s==0, (ix-kx)%s==0} dy[n][(iy-ky)/s][(ix-kx)/s][o] * w[o][c][ky][kx]
 * dy: NHWC fp32 [n][ho][wo][co] (no halo); wg: mode-2 pack (f32); dx: NHWC fp32 [n][hp][wp][ci]
 * (gradient w.r.t. the padded operand).  SIMT kernel; stride-1 layers use skit_conv2d_fwd with a
 * mode-1 pack on the tensor-core path instead. */
int skit_conv2d_dgrad_gather(const float* dy, int n, int ho, int wo, int co,
                             const skit_weights* wg, int stride, int hp, int wp, float* dx, void* stream);

/* Precision of the BACKWARD tensor-core launches (skit_conv2d_dgrad_s1 / _s2, skit_conv2d_wgrad*): the autograd of F.conv2d
 * the reference runs in TF32 on its own GPUs (torch.backends.cudnn.allow_tf32 default, README.md:61 torch 1.11).
 *   terms = 3: every bf16 hi/lo pair is multiplied as A_lo*B_hi + A_hi*B_lo + A_hi*B_hi (like the forward convs, ~1e-5 of fp32);
 *   terms = 2 (default; SKIT_BWD_TERMS=3 in the environment flips it): the filter's lo plane (dgrad) / the output gradient's lo
 *              plane (wgrad) is neither loaded nor multiplied — gradients within ~5e-3 of fp32 (gate: 3e-2 per tensor).
 * Forward launches (skit_conv2d_fwd) are always three-term. */
int skit_set_backward_terms(int terms);

/* Input gradient of a stride-1 conv (autograd of F.conv2d, e.g. the ResnetBlock convs networks.py:1281-1310):
 *   dx[n][y][x][c] = sum_{a,b,o} dz[n][y+a][x+b][o] * w[o][c][k-1-a][k-1-b],  dz = dy with a zero halo of k-1
 * dy: operand carrying that halo (dy->hp = H + k - 1 for an H x W gradient w.r.t. the conv's padded input); w1: mode-1 pack;
 * dx: NHWC fp32 [n][dy->hp - k + 1][dy->wp - k + 1][w1->co].  On the halo-tile tcgen05 kernel the one-pixel border strips
 * (whose windows touch the data through a single filter row / column) run as extra regions of the same grid. */
int skit_conv2d_dgrad_s1(const skit_operand* dy, const skit_weights* w1, float* dx, void* stream);

/* Input gradient of a stride-2, even-k conv on the tensor cores (PatchGAN k4 s2 layers, networks.py:1706-1716):
 * the four parities (iy%2, ix%2) of dx are four stride-1 (k/2 x k/2) convolutions of dy — which must be a
 * bf16x2 operand zero-haloed by dy_pad = k/2-1 — with the mode-3 sub-filters; each writes its interleaved
 * quarter of dx (NHWC fp32 [n][hp][wp][ci], every element written exactly once). */
int skit_conv2d_dgrad_s2(const skit_operand* dy, int dy_pad, const skit_weights* wp, int k,
                         int ho, int wo, int hp, int wp_, float* dx, void* stream);

/* nn.ConvTranspose2d forward (thirdparty/unet/unet_parts_custom.py:63: k4, stride 2, padding 1), as the gather form
 * of a strided conv's input gradient cropped by `pad`:
 *   y[n][oy][ox][y_c0 + c] = bias[c] + sum_{ky,kx,o : (oy+pad-ky)%s==0, (ox+pad-kx)%s==0} x[n][(oy+pad-ky)/s][(ox+pad-kx)/s][o] * w[o][c][ky][kx]
 * x: dense NHWC fp32 [n][h][w][ci]; wg: mode-2 pack of the ConvTranspose2d weight [ci][co][k][k] taken as a conv weight
 * (co_conv = ci, ci_conv = co); y: NHWC fp32 with y_ctot channels, this call fills [y_c0, y_c0 + co).  stats as in
 * skit_conv2d_fwd.  Its backward is skit_conv2d_fwd (dx: stride-s conv of the pad-haloed dy with the mode-0 pack) and
 * skit_conv2d_wgrad with the operand roles swapped (x := haloed dy, dy := x). */
int skit_conv_transpose2d_fwd(const float* x, int n, int h, int w, int ci, const skit_weights* wg, int stride, int pad,
                              int ho, int wo, const float* bias, float* y, int y_ctot, int y_c0,
                              double* stats, int stats_mode, void* stream);

/* dbias[o] += sum over images and the [ho][wo] window at dy_org of an operand (bias gradient of a layer without norm). */
int skit_dbias(const skit_operand* dy, int dy_org, int ho, int wo, float* dbias, void* stream);

/* Weight gradient (autograd of F.conv2d w.r.t. weight and bias):
 *   dw[o][c][ky][kx] += sum_{n,oy,ox} dy[n][oy][ox][o] * x[n][org+oy*s+ky][org+ox*s+kx][c]
 * dy is an operand (fp32 or bf16x2) read with halo offset dy_org.  `scratch` is k*k*ci*co floats, ZEROED by
 * the caller: the split-K partial sums land there (atomics) before being folded into dw (reference layout,
 * accumulated).  dbias (may be NULL): [co] += sum dy.  tcgen05 path when both operands are bf16x2, channel
 * counts are multiples of 64 and one of them of 128 (both GEMM operands MN-major straight from NHWC). */
int skit_conv2d_wgrad(const skit_operand* x, int org, const skit_operand* dy, int dy_org,
                      int k, int stride, int ho, int wo, float* scratch, float* dw, float* dbias, int impl, void* stream);

/* Same, for channel-padded operands: x->c >= ci_real and dy->c >= co_real (padding channels are zero); scratch is
 * k*k*x->c*dy->c floats; dw has the real shape [co_real][ci_real][k][k]; dbias (may be NULL): [co_real]. */
int skit_conv2d_wgrad_ex(const skit_operand* x, int org, const skit_operand* dy, int dy_org,
                         int k, int stride, int ho, int wo, float* scratch, float* dw, float* dbias, int impl,
                         int co_real, int ci_real, void* stream);

/* ---------------------------------------------------------------- normalisation + activation + halo
 * Turn double (sum, sumsq) into float (mean, rstd), eps 1e-5, biased variance
 * (nn.InstanceNorm2d networks.py:138-139; nn.BatchNorm2d training mode networks.py:1714-1723).
 * For BatchNorm also updates running_mean / running_var (momentum, unbiased var) when non-NULL. */
int skit_stats_finalize(const double* stats, int groups, int c, double count, float eps,
                        float* mean_rstd, float* running_mean, float* running_var, float momentum, void* stream);

/* y = act(norm(raw) * gamma + beta) [+ residual]  →  optional dense copy `out` (NHWC fp32) and/or a
 * haloed operand `op` (pad, pad_mode, fmt).  Fuses InstanceNorm/BatchNorm-apply, ReLU/LeakyReLU,
 * the ResnetBlock skip add (networks.py:1322), ReflectionPad2d / zero padding and the bf16 hi/lo
 * split for the tensor-core conv that follows.  mean_rstd: [groups][c][2] or NULL (norm none). */
int skit_norm_act_pad(const float* raw, int n, int h, int w, int c,
                      const float* mean_rstd, int norm_mode, const float* gamma, const float* beta,
                      int act, const float* residual, float* out,
                      const skit_operand* op, int pad, int pad_mode, void* stream);

/* skit_norm_act_pad_ex with the statistics finalise fused in (InstanceNorm2d / BatchNorm2d of models/networks.py:138-139 and
 * :1712-1723 between a conv and the next one): `stats` are the fp64 [groups][c][2] (sum, sum of squares) the producing conv's
 * epilogue accumulated over `count` elements; every thread block derives mean and 1/sqrt(var + eps) for its own channels (same
 * arithmetic as skit_stats_finalize) and `mean_rstd_out` [groups][c][2] is written as a side output for the backward pass —
 * no separate finalise launch sits between a conv and the operand of the next one. */
int skit_norm_act_pad_stats(const float* raw, int n, int h, int w, int c,
                            const double* stats, double count, float eps, float* mean_rstd_out,
                            int norm_mode, const float* gamma, const float* beta,
                            int act, const float* residual, float* out,
                            const skit_operand* op, int c_off, int pad, int pad_mode, void* stream);

/* Same, writing channels [c_off, c_off + c) of a wider operand (op->c >= c_off + c): builds `ReLU(cat(x, skip))`
 * of the U-Net decoder (thirdparty/unet/unet_parts_custom.py:46-79) slice by slice, never materialising the cat. */
int skit_norm_act_pad_ex(const float* raw, int n, int h, int w, int c,
                         const float* mean_rstd, int norm_mode, const float* gamma, const float* beta,
                         int act, const float* residual, float* out,
                         const skit_operand* op, int c_off, int pad, int pad_mode, void* stream);

/* Backward, phase A: fold the halo gradient back (adjoint of the padding), add an optional dense
 * gradient, apply act', and reduce the two norm-backward sums.
 *   g = act'(pre) * ( fold(dpad) + dadd ),  pre = norm(raw)*gamma+beta
 *   sums[r][group][c][0] += sum g ; sums[r][group][c][1] += sum g * xhat   (double; r = one of SKIT_SUM_REPLICAS copies)
 * dpad: NHWC fp32 [n][h+2pad][w+2pad][c] or NULL; dadd: NHWC fp32 [n][h][w][c] or NULL. */
int skit_act_norm_bwd_reduce(const float* dpad, int pad, int pad_mode, const float* dadd,
                             const float* raw, int n, int h, int w, int c,
                             const float* mean_rstd, int norm_mode, const float* gamma, const float* beta,
                             int act, float* g, double* sums, void* stream);

/* Same with the dense gradient(s) given as channel slices [dadd_c0, dadd_c0 + c) of NHWC tensors with dadd_ctot
 * channels (the gradient of a concatenated decoder input), an optional second one (twin RGB / touch decoders,
 * networks.py:1635-1644), and an optional ReLU mask [pre > 0] on them (the decoder's in-place ReLU acts on the
 * concat, i.e. on the LeakyReLU'd skip tensor: SURVEY.md section 3.3). */
int skit_act_norm_bwd_reduce_ex(const float* dpad, int pad, int pad_mode,
                                const float* dadd, const float* dadd2, int dadd_c0, int dadd_ctot, int dadd_relu_mask,
                                const float* raw, int n, int h, int w, int c,
                                const float* mean_rstd, int norm_mode, const float* gamma, const float* beta,
                                int act, float* g, double* sums, void* stream);

/* Same, with an optional second output dsum [n][h][w][c] = fold(dpad) + dadd (+ dadd2): the incoming gradient itself, before
 * act'.  A ResnetBlock's output gradient feeds both its conv branch and the skip path (networks.py:1322): one pass writes both. */
int skit_act_norm_bwd_reduce_ex2(const float* dpad, int pad, int pad_mode,
                                 const float* dadd, const float* dadd2, int dadd_c0, int dadd_ctot, int dadd_relu_mask,
                                 const float* raw, int n, int h, int w, int c,
                                 const float* mean_rstd, int norm_mode, const float* gamma, const float* beta,
                                 int act, float* g, double* sums, float* dsum, void* stream);

/* Backward, phase B: d_raw = gamma*rstd*( g - S0/count - xhat*S1/count )  (norm none: d_raw = g),
 * written as a zero-haloed operand (pad q) in fmt, ready for dgrad/wgrad.
 * For BatchNorm, dgamma[c] += S1, dbeta[c] += S0 when non-NULL. */
int skit_norm_bwd_apply(const float* g, const float* raw, int n, int h, int w, int c,
                        const float* mean_rstd, int norm_mode, const float* gamma,
                        const double* sums, double count, float* dgamma, float* dbeta,
                        const skit_operand* op, int pad, void* stream);

/* Same, with an optional dense gradient `extra` [n][h][w][c] added to d_raw: the gradient of a feature tapped on the raw
 * conv output (ResnetGenerator.forward(layers=...), networks.py:1131-1150, as used by the PatchNCE wiring). */
int skit_norm_bwd_apply_ex(const float* g, const float* raw, int n, int h, int w, int c,
                           const float* mean_rstd, int norm_mode, const float* gamma,
                           const double* sums, double count, float* dgamma, float* dbeta,
                           const float* extra, const skit_operand* op, int pad, void* stream);

/* Antialiased resamplers (networks.py:51-74 Downsample, :87-107 Upsample), NHWC fp32 dense.
 * down: reflect pad 1, depthwise [1,2,1]^2/16, stride 2  ([h][w] -> [h/2][w/2], even h,w)
 * up:   replicate pad 1, depthwise conv_transpose [1,3,3,1]^2*4/64 stride 2, cropped -> [2h][2w] */
int skit_blur_down_fwd(const float* x, int n, int h, int w, int c, float* y, void* stream);
int skit_blur_down_bwd(const float* dy, int n, int h, int w, int c, float* dx, void* stream);
int skit_blur_up_fwd(const float* x, int n, int h, int w, int c, float* y, void* stream);
int skit_blur_up_bwd(const float* dy, int n, int h, int w, int c, float* dx, void* stream);

/* ---------------------------------------------------------------- image-level ops (NCHW fp32 planar)
 * Concatenate up to 4 NCHW sources along channels into a haloed NHWC fp32 operand
 * (torch.cat + ReflectionPad2d(3) at networks.py:1077 / zero padding 2 of the PatchGAN convs :1703).
 * The operand may be fp32 or bf16x2 and may have MORE channels than the sources provide (zero channel padding). */
int skit_nchw_cat_to_operand(const float* const* srcs, const int* chans, int nsrc,
                             int n, int h, int w, const skit_operand* op, int pad, int pad_mode, void* stream);
/* Adjoint for one channel slice: dst[n][cs][h][w] (+)= fold(dpad)[.., c0:c0+cs]. */
int skit_operand_grad_to_nchw(const float* dpad, int n, int h, int w, int c, int pad, int pad_mode,
                              int c0, int cs, float* dst, int accumulate, void* stream);

/* Generator head: raw [n][h][w][5] (conv + bias) -> tanh -> *M -> fake_I [n][3][h][w], fake_T [n][2][h][w],
 * fake_N = normalize([gx, gy, scale_nz]) (networks.py:1127; sinskitG_model.py:1309-1319; model_utils.py:418-425). */
int skit_g_head_fwd(const float* raw, const float* mask, int n, int h, int w, float scale_nz,
                    float* fake_I, float* fake_T, float* fake_N, void* stream);
/* d_raw[n][h][w][5] = [dI, dT] * M * (1 - tanh(raw)^2), written as a zero-haloed operand (pad q), fp32 or bf16x2,
 * with op->c >= 5 channels (zero channel padding for the tensor-core input-gradient / weight-gradient kernels). */
int skit_g_head_bwd(const float* raw, const float* mask, const float* dI, const float* dT,
                    int n, int h, int w, const skit_operand* op, int pad, void* stream);

/* Same gradient as two zero-haloed fp32 operands (3 RGB channels, 2 touch channels) for the twin decoders of the
 * default U-Net generator, whose last layers are separate transposed convs (networks.py:1635-1644). raw: [n][h][w][5]. */
int skit_g_head_bwd_split(const float* raw, const float* mask, const float* dI, const float* dT,
                          int n, int h, int w, const skit_operand* opI, const skit_operand* opT, int pad, void* stream);

/* x[n][c][h][w] *= m[n][0][h][w] in place: the `S *= M`, `I *= M`, `T *= I_masks` of set_input
 * (sinskitG_model.py:724,734,789-790) on the device, after the pinned H2D copy of the raw tensors. */
int skit_mask_mul(float* x, const float* m, int n, int c, int h, int w, void* stream);

/* y[n][0][h][w] = mean_c x[n][c][h][w] and its adjoint dx[n][c][h][w] += dy[n][0][h][w] / c  (NCHW planes): the
 * 1-channel view of the generated image that the PatchNCE query branch re-encodes (PatchNCE wiring). */
int skit_channel_mean(const float* x, int n, int c, int h, int w, float* y, void* stream);
int skit_channel_mean_bwd(const float* dy, int n, int c, int h, int w, float* dx, void* stream);

/* DiffAugment policy 'bs' followed by *M (thirdparty/DiffAugment.py:25-33; sinskitG_model.py:1330-1340).
 * u_b, u_s: [n] device floats (the host's torch.rand draws). x: [n][3][h][w]. */
int skit_diffaug_bs_mask(const float* x, const float* mask, const float* u_b, const float* u_s,
                         int n, int h, int w, float* y, void* stream);

/* AvgPool2d(3, stride 2, padding 1, count_include_pad=False) (networks.py:1670), NCHW planes. */
int skit_avgpool3s2_fwd(const float* x, int planes, int h, int w, float* y, void* stream);
int skit_avgpool3s2_bwd(const float* dy, int planes, int h, int w, float* dx, int accumulate, void* stream);

/* Patch gather (get_patch_in_input, model_utils.py:254-335): for each of np patches and each source
 * s, dst[p][coff_s + c][y][x] = src_s[0][c][clamp(oy[p]+y)][clamp(ox[p]+x)]; dst is NCHW
 * [np][ctot][ps][ps].  ox/oy: int32 device arrays.  One launch covers all sources. */
int skit_patch_gather(const float* const* srcs, const int* chans, const int* coffs, int nsrc,
                      int h, int w, const int* ox, const int* oy, int np, int ps,
                      float* dst, int ctot, void* stream);
/* Adjoint: dsrc[0][c][clamp(..)][clamp(..)] += dpatch[p][coff + c][y][x]  (atomic scatter-add). */
int skit_patch_scatter_add(const float* dpatch, int ctot, int coff, int cs, int h, int w,
                           const int* ox, const int* oy, int np, int ps, float* dsrc, void* stream);

/* GANLoss 'nonsaturating' on one scale (networks.py:500-522): loss[b] += mean_hw softplus(sign * pred[b]).
 * sign = -1 for target_is_real, +1 for fake.  If dpred != NULL: dpred = gscale * sign * sigmoid(sign*pred)/(h*w). */
int skit_gan_softplus(const float* pred, int n, int hw, float sign, float* loss, float* dpred, float gscale, void* stream);
/* GANLoss.get_loss_for_single_scale_discriminator in every mode (models/networks.py:500-522): mode 0 nonsaturating, 1 hinge, 2 wgan /
 * wgangp (value only: the gradient penalty is not part of GANLoss), 3 lsgan (MSE against `target`), 4 vanilla (BCE-with-logits against
 * `target`).  loss[b] += mean over sample b's hw elements; dpred (optional) = gscale * d(that mean)/d(pred). */
int skit_gan_loss(const float* pred, int n, int hw, int mode, int target_is_real, float target, float* loss, float* dpred,
                  float gscale, void* stream);

/* L1: loss[0] += scale * sum |a-b| ; if grad != NULL: grad (+)= gscale * sign(a-b)  (sinskitG_model.py:1702,1812). */
int skit_l1_loss(const float* a, const float* b, long long numel, float scale, float* loss,
                 float* grad, float gscale, int accumulate, void* stream);

/* ---------------------------------------------------------------- optimiser
 * torch.optim.Adam step on a flat fp32 bucket (sinskitG_model.py:590-599), optional grad scaling
 * (1/world_size after the all-reduce). step >= 1. */
int skit_adam_step(float* p, const float* g, float* m, float* v, long long numel, int step,
                   float lr, float beta1, float beta2, float eps, float grad_scale, void* stream);

/* Same update, with the two step-dependent scalars in DEVICE memory: hyper = [lr / (1 - beta1^step),
 * 1 / sqrt(1 - beta2^step)] (fp32).  Lets one captured CUDA graph of the whole train step be replayed
 * while the step count and the LambdaLR factor (networks.py:161-165) advance. */
int skit_adam_step_dev(float* p, const float* g, float* m, float* v, long long numel, const float* hyper,
                       float beta1, float beta2, float eps, float grad_scale, void* stream);

/* ---------------------------------------------------------------- PatchNCE
 * PatchSampleF gather + Normalize (networks.py:689-719, 585-594): feat NHWC fp32 [b][hw][c];
 * out[b*np + i][c] = f[b][ids[i]][c] / (||f||_2 + 1e-7).  `pre` (optional) keeps the un-normalised rows. */
int skit_patch_sample_l2norm(const float* feat, int b, int hw, int c, const int* ids, int np,
                             float* out, float* pre, void* stream);
/* Backward of the normalisation + gather: dfeat[b][ids[i]][c] += d(pre) */
int skit_patch_sample_l2norm_bwd(const float* dout, const float* pre, int b, int hw, int c,
                                 const int* ids, int np, float* dfeat, void* stream);
/* Adjoint of the plain row gather (PatchSampleF with an MLP, networks.py:706-712: the MLP sits between the gather and
 * the normalisation): dfeat[b][ids[i]][c] += drows[b*np + i][c]. */
int skit_rows_scatter_add(const float* drows, int b, int hw, int c, const int* ids, int np, float* dfeat, void* stream);
/* PatchNCELoss.forward (patchnce.py:13-55): q,k [b*np][dim]; loss[b*np]; optional dq (gradient of
 * sum_i loss[i]*gscale w.r.t. q; k is detached in the reference). */
int skit_patchnce(const float* q, const float* k, int b, int np, int dim, float inv_T,
                  float* loss, float* dq, float gscale, void* stream);

/* ---------------------------------------------------------------- profiling aid
 * When buf != NULL the halo-tile conv kernel stores clock64() stamps per CTA into buf[cta][8]
 * (start, setup done, first activation tile landed, last MMA issued, accumulator ready, stores done, end).
 * Pass NULL to switch it off (the default).  Never enabled on the product path. */
int skit_debug_set_buffer(long long* buf);

/* ---- StyleGAN2 generator (models/stylegan_networks.py; `--netG stylegan2 | smallstylegan2`) ------------------------------
 * Effective conv filters, rebuilt on the device whenever the parameters change.  w: reference layout [co][ci][k][k]
 * (ModulatedConv2d's [1][co][ci][k][k] is the same memory).
 *   mode 0  EqualConv2d (:179-190):        out[co][ci][k][k]   = w / sqrt(ci k k)
 *   mode 1  Blur + stride-2 conv (:625-643) folded into one filter: out[co][ci][k+3][k+3] = (w / sqrt(ci k k)) (*) blur4x4,
 *           run as a stride-2 conv with zero padding 2 (k = 3) or 1 (k = 1, the ResBlock's blur + 1x1 stride-2 skip, :677).
 *   mode 2  ModulatedConv2d(upsample, style = 1) (:304-334): weights demodulated per output channel, then
 *           conv_transpose2d(stride 2) + Blur(pad 1,1, x4) folded into four 3x3 pad-1 filters, one per output parity:
 *           out[(py*2+px)*co + o][ci][3][3]; the conv's 4 co output channels are the 2x2 sub-pixels (depth-to-space). */
int skit_sg2_weight_prep(const float* w, int co, int ci, int k, int mode, float* out, void* stream);
/* FusedLeakyReLU / NoiseInjection / ResBlock merge (:18-35, :351-362, :683-689) in one pass over a conv's raw output:
 *   out = (lrelu_0.2(raw + bias + *noise_w * noise) * gain + skip) * post      (act = 0: no lrelu / gain)
 * raw [n][h][w][craw]; shuffle = 1 reads it as 2x2 sub-pixel groups of c channels (mode-2 filters) and writes 2h x 2w;
 * skip: dense [n][H][W][c].  noise [n][H][W].  Outputs (any subset): dense NHWC fp32,
 * a zero-haloed operand (fp32 or bf16x2) for the next conv, NCHW fp32 with the first nchw_c channels.  c % 4 == 0. */
int skit_sg2_bias_act(const float* raw, int n, int h, int w, int craw, int c, const float* bias,
                      const float* noise, const float* noise_w, const float* skip, int shuffle,
                      int act, float gain, float post, float* dense, const skit_operand* op, int pad, float* nchw,
                      int nchw_c, void* stream);
/* Backward of skit_sg2_bias_act.  The gradient w.r.t. its OUTPUT arrives as up to two zero-haloed tensors
 * ([n][H+2p][W+2p][c] fp32: input gradients of the convs that consumed it) and / or a dense one; d_out is their sum.
 *   draw  (operand [n][h+2q][w+2q][c or 4c], zero halo, fp32 or bf16x2) = d_out * post * (act ? gain * lrelu'(raw + bias + nw noise) : 1),
 *         written in the RAW conv output's layout (space-to-depth when shuffle) — the conv's backward operand;
 *   dskip (optional, dense) = d_out * post;  dbias[c] += sum draw;  *dnoise_w += sum draw * noise. */
int skit_sg2_bias_act_bwd(const float* raw, int n, int h, int w, int craw, int c, const float* bias,
                          const float* noise, const float* noise_w, int shuffle, int act, float gain, float post,
                          const float* dpad_a, int pad_a, const float* dpad_b, int pad_b, const float* ddense,
                          const skit_operand* draw, int q, float* dskip, float* dbias, float* dnoise_w, void* stream);
/* Transpose of skit_sg2_weight_prep: dw[co][ci][k][k] += (d eff / d w)^T deff, through the blur fold, the EqualConv scale and
 * (mode 2) the demodulation. */
int skit_sg2_weight_prep_bwd(const float* w, int co, int ci, int k, int mode, const float* deff, float* dw, void* stream);

/* ---- LPIPS-VGG16 perceptual loss (pip `lpips` 0.1.4 LPIPS(net='vgg'); call sites models/sinskitG_model.py:495, 1639-1645, 1711)
 * The VGG16 trunk runs on skit_conv2d_fwd / skit_conv2d_dgrad_* with frozen packs; these are the pieces around it.
 * ScalingLayer: op[n][h+2][w+2][3] (fp32, zero halo) = (x - shift) / scale, x NCHW with 1 (broadcast) or 3 channels. */
int skit_lpips_scale_fwd(const float* x, int n, int cin, int h, int w, const skit_operand* op, void* stream);
/* its transpose: dx[n][dx_c0 .. dx_c0+cin)[h][w] of an [n][dx_ctot][h][w] tensor (+)= gscale * dop / scale (3 channels summed when cin = 1) */
int skit_lpips_scale_bwd(const float* dop, int n, int cin, int h, int w, float gscale, float* dx, int dx_ctot, int dx_c0,
                         int accumulate, void* stream);
/* MaxPool2d(2, 2) of a dense NHWC map into the next conv's zero-haloed operand (fp32 or bf16x2), and its backward:
 * df = add (optional, dense) + dpool routed to the first maximum of each window (ATen's tie rule); dpool is the gradient
 * w.r.t. the pooled haloed operand [n][h/2+2p][w/2+2p][c]. */
int skit_maxpool2_fwd(const float* f, int n, int h, int w, int c, const skit_operand* op, int pad, void* stream);
int skit_maxpool2_bwd(const float* f, const float* dpool, int n, int h, int w, int c, int pad, const float* add, float* df, void* stream);
/* One LPIPS layer: loss[b] += mean_pixels sum_c lin_w[c] (u0 - u1)^2, u = f / (|f|_2 over channels + 1e-10);
 * df0 (optional) = gscale * d loss[b] / d f0.  f0, f1: [n][h][w][c] fp32. */
int skit_lpips_layer(const float* f0, const float* f1, const float* lin_w, int n, int h, int w, int c, float gscale,
                     float* loss, float* df0, void* stream);

/* Zero-fill `nbytes` at p on `stream` with cudaMemsetAsync (a memset node when the stream is being captured): the zero_grad of
 * the flat gradient buckets (optimizer.zero_grad, sinskitG_model.py:648-694) and the scratch / scatter targets of the backward pass. */
int skit_zero_bytes(void* p, long long nbytes, void* stream);

/* ---- Evaluation metrics (models/model_utils.py:431-561 compute_evaluation_metric).  Device-side
 * reductions into fp64 / fp32 scalars the caller zero-initialises (minmax: {+inf, -inf}); nothing synchronises.
 *   skit_metric_minmax        out2 = {min(x), max(x)}: the real image's range for the [0, 1] rescale (:483-487)
 *   skit_metric_sq_err        out += sum (a' - b')^2;  minmax != NULL: a' = (a - lo)/(hi - lo), b' = clamp((b - lo)/(hi - lo), 0, 1)
 *                             (PSNR, :494-495); else clamp_b: b' = clamp(b, 0, 1) (T_MSE, :520,553-555)
 *   skit_metric_ssim          out += sum of the SSIM index map (11x11 Gaussian, sigma 1.5, k1 0.01, k2 0.03, reflect pad 5, border of 5
 *                             dropped: torchmetrics functional/image/ssim.py, the pip package behind `SSIM` at :497-498) over `planes`
 *                             contiguous h x w planes; same rescale as above when minmax is given
 *   skit_metric_normal_angle  out += sum over pixels of the angle in degrees between F.normalize([gx, gy, scale_nz]) of the two touch
 *                             maps [n][2][h][w] (model_utils.py:418-425 + models/normal_losses.py:10-33, mode 'evaluate') */
int skit_metric_minmax(const float* x, long long n, float* out2, void* stream);
int skit_metric_sq_err(const float* a, const float* b, long long n, const float* minmax, int clamp_b, double* out, void* stream);
int skit_metric_ssim(const float* a, const float* b, int planes, int h, int w, const float* minmax, float data_range,
                     double* out, void* stream);
int skit_metric_normal_angle(const float* real, const float* fake, int n, int h, int w, float scale_nz, int clamp_fake,
                             double* out, void* stream);

/* ---- Data pipeline (data/singleskit_dataset.py, data/dataset_util.py).  Byte / integer / fp64 work,
 * bit-identical with what the reference computes on the host through Pillow, NumPy and OpenCV.  Images are HWC uint8 device buffers.
 *   skit_resize_u8            `PIL.Image.resize((dw, dh), filter)` for 8-bit images of 1..4 bands (dataset_util.py:159-163 zoom_img,
 *                             :194-201 crop_img, :219-231 make_power_2_img): Pillow's two-pass fixed-point resampler
 *                             (src/libImaging/Resample.c; coefficients in double on the host, 22-bit fixed point on the device).
 *                             `filter` takes PIL.Image.Resampling's own values.
 *   skit_u8_crop_to_tensor    `img.crop((x0, y0, x0+w, y0+h))` (pixels outside read 0) + `transforms.ToTensor()` [+ `Normalize(0.5, 0.5)`]
 *                             (singleskit_dataset.py:301-315): fp32 CHW.
 *   skit_contact_centers      the centre loop of process_all_valid_patches (singleskit_dataset.py:745,768-803) for P touch patches stored
 *                             back to back (pix_off[P+1] pixel offsets): for each patch, in_mask = whether its rectangle touches the object
 *                             mask M, and the row-major list (as np.where orders it) of the linear pixel indices with center_mask > 0 whose
 *                             patch x patch window of touch_mask * M_patch / 255 reaches 1; centers uses the same per-patch offsets,
 *                             counts[p] entries are valid.  scratch = 3 * total_pixels bytes.
 *   skit_touch_squares        the sampled squares (:812-860): for K selections (patch, cx, cy) the patch x patch windows of gx and gy
 *                             (elements of elem_size 4 or 8 bytes, copied verbatim) -> t_images [K][2][patch][patch], and
 *                             square_mask = touch_mask * M_patch / 255 in fp64 -> i_masks [K][patch][patch].
 *   skit_laplacian_var_u8     `variance_of_laplacian(S_patch, ref=255)` (util/util.py:261-265) of K size x size windows at (x0, y0) of a
 *                             single-band image (outside reads 0, as PIL's crop): uint8 wrap of (image - ref), cv2.Laplacian ksize 1 with
 *                             reflect-101 borders, population variance in fp64 = the resampling weight before its clamp (:1053-1056). */
#define SKIT_RESAMPLE_LANCZOS 1
#define SKIT_RESAMPLE_BILINEAR 2
#define SKIT_RESAMPLE_BICUBIC 3
#define SKIT_RESAMPLE_BOX 4
#define SKIT_RESAMPLE_HAMMING 5
int skit_resize_u8(const unsigned char* src, int sh, int sw, int c, unsigned char* dst, int dh, int dw, int filter, void* stream);
int skit_u8_crop_to_tensor(const unsigned char* src, int sh, int sw, int c, int y0, int x0, int h, int w, int normalize,
                           float* dst, void* stream);
int skit_contact_centers(const double* touch_mask, const unsigned char* center_mask, const long long* pix_off, long long total_pixels,
                         const int* ph, const int* pw, const int* roi_x, const int* roi_y, int P, const unsigned char* M, int mh,
                         int mw, int patch, unsigned char* scratch, int* counts, int* centers, int* in_mask, void* stream);
int skit_touch_squares(const double* touch_mask, const long long* pix_off, const int* ph, const int* pw, const int* roi_x,
                       const int* roi_y, int P, const unsigned char* M, int mh, int mw, const void* gx, const void* gy, int elem_size,
                       const int* sel_patch, const int* sel_cx, const int* sel_cy, int K, int patch, void* t_images, double* i_masks,
                       void* stream);
int skit_laplacian_var_u8(const unsigned char* img, int h, int w, const int* x0, const int* y0, int K, int size, int ref, double* out,
                          void* stream);

/* Candidate offsets of get_patch_in_input's random mode (models/model_utils.py:212-218): clamp(conv2d(M, ones(k, k), padding=pad), 0, 1)
 * != 0 for a non-negative single-channel mask M [h][w], on the oh x ow = (h + 2 pad - k + 1) x (w + 2 pad - k + 1) map.  bits: oh rows of
 * ceil(ow / 32) 32-bit words (bit c % 32 of word c / 32 = position (r, c), the order torch.nonzero lists them in); rowcount[oh] = set
 * bits per row.  scratch = h * ow bytes. */
int skit_mask_box_bits(const float* M, int h, int w, int k, int pad, unsigned char* scratch, unsigned int* bits, int* rowcount,
                       void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SKIT_B200_H */
