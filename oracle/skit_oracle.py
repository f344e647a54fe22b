"""TEST INFRASTRUCTURE ONLY — CPU fp32 restatement of the reference's skitG/sinskitG hot path.

This module is the *oracle* (parity checker) for the CUDA path in
`visual-tactile-synthesis_b200/`.  It is imported only by tests/, `__graft_entry__.smoke()`
and bench.py's `cpu_baseline` / `--impl reference` legs.  The product never routes
through it.

Pinning: the reference ships no golden vectors for this path, so the
oracle is pinned against *outputs of the real reference code* run in the build container:
`oracle/make_golden.py` imports /root/reference (via oracle/ref_loader.py), runs its
`define_G/define_D/GANLoss/get_patch_in_input/DiffAugment/SPE/PatchNCELoss/PatchSampleF` and
the full `SinSKITGModel.optimize_parameters`, and commits the results under tests/golden/.
tests/test_oracle_golden.py checks every function here against those fixtures.

Exception: `lpips_vgg` restates the third-party pip package lpips 0.1.4 (absent from /root/reference and from this image);
its VGG16 trunk is pinned against torchvision's vgg16 module, its head is PARITY UNPINNED (no fixture exists to check it);
the train step's USE of it is pinned against the real reference's call sites (tests/golden/step_resnet_lpips.npz).
The StyleGAN2 generator restatement is pinned by tests/golden/stylegan2.npz (bit-identical to the reference on CPU).

Everything is a pure function over a flat `state` dict {reference state_dict key: tensor}
(fp32, NCHW, reference shapes), so the same weights drive the oracle and the CUDA path.
Each function cites the reference file:line it restates (paths relative to /root/reference).
"""
import math
import random as _pyrandom

import numpy as np
import torch
import torch.nn.functional as F

EPS_NORM = 1e-5


# --------------------------------------------------------------------------- positional enc.
def spe_grid(h, w, emb_dim=4, n=1):
    """SinusoidalPositionalEmbedding(emb_dim, padding_idx=0).make_grid2d
    (thirdparty/mmgeneration/positional_encoding.py:61-86, 113-159).
    Channels: [sin(p*f_0..), cos(p*f_0..)] of the x position (1..w), then of the y position."""
    half = emb_dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000.0) / (half - 1)))

    def table(length):
        pos = torch.arange(1, length + 1, dtype=torch.float32)[:, None] * freq[None, :]
        return torch.cat([torch.sin(pos), torch.cos(pos)], dim=1)  # [length, emb_dim]

    ex = table(w).t()[None, :, None, :].expand(n, emb_dim, h, w)
    ey = table(h).t()[None, :, :, None].expand(n, emb_dim, h, w)
    return torch.cat([ex, ey], dim=1).contiguous()


# --------------------------------------------------------------------------- small helpers
def inorm(x):
    """nn.InstanceNorm2d(affine=False, track_running_stats=False) (models/networks.py:138-139)."""
    mu = x.mean(dim=(2, 3), keepdim=True)
    var = x.var(dim=(2, 3), unbiased=False, keepdim=True)
    return (x - mu) / torch.sqrt(var + EPS_NORM)


_BLUR3 = torch.tensor([1.0, 2.0, 1.0])
_BLUR4 = torch.tensor([1.0, 3.0, 3.0, 1.0])


def blur_down(x):
    """Downsample(C): reflect pad 1, depthwise [1,2,1]^2/16, stride 2 (models/networks.py:51-74)."""
    c = x.shape[1]
    k = (_BLUR3[:, None] * _BLUR3[None, :])
    k = (k / k.sum()).to(x)[None, None].repeat(c, 1, 1, 1)
    return F.conv2d(F.pad(x, (1, 1, 1, 1), mode="reflect"), k, stride=2, groups=c)


def blur_up(x):
    """Upsample(C): replicate pad 1, depthwise conv_transpose [1,3,3,1]^2*4/64, stride 2,
    padding 2, drop first row/col and (even filter) last row/col (models/networks.py:87-107)."""
    c = x.shape[1]
    k = (_BLUR4[:, None] * _BLUR4[None, :])
    k = (k / k.sum() * 4.0).to(x)[None, None].repeat(c, 1, 1, 1)
    y = F.conv_transpose2d(F.pad(x, (1, 1, 1, 1), mode="replicate"), k, stride=2, padding=2, groups=c)
    return y[:, :, 1:-1, 1:-1]


# --------------------------------------------------------------------------- generators
def resnet_g_forward(sd, x, n_blocks=9, layers=(), encode_only=False):
    """ResnetGenerator.forward with the antialiased down/up path
    (models/networks.py:1077-1154; ResnetBlock :1293-1324; index map SURVEY.md A.2).
    `layers` taps the output of nn.Sequential index i, as the reference's `layers=` does."""
    feats = []
    idx = [0]
    want = set(layers)
    last = max(layers) if len(layers) else -1

    class _Done(Exception):
        pass

    def tap(t):
        if idx[0] in want:
            feats.append(t)
        if encode_only and idx[0] == last:
            raise _Done()
        idx[0] += 1
        return t

    def conv(t, i, pad=0):
        return F.conv2d(t, sd["model.%d.weight" % i], sd.get("model.%d.bias" % i), padding=pad)

    try:
        t = tap(F.pad(x, (3, 3, 3, 3), mode="reflect"))          # 0
        t = tap(conv(t, 1))                                       # 1
        t = tap(inorm(t))                                         # 2
        t = tap(F.relu(t))                                        # 3
        t = tap(conv(t, 4, 1)); t = tap(inorm(t)); t = tap(F.relu(t)); t = tap(blur_down(t))    # 4-7
        t = tap(conv(t, 8, 1)); t = tap(inorm(t)); t = tap(F.relu(t)); t = tap(blur_down(t))    # 8-11
        m = 12
        for b in range(n_blocks):
            p = "model.%d.conv_block." % (m + b)
            r = F.pad(t, (1, 1, 1, 1), mode="reflect")
            r = F.conv2d(r, sd[p + "1.weight"], sd.get(p + "1.bias"))
            r = F.relu(inorm(r))
            r = F.pad(r, (1, 1, 1, 1), mode="reflect")
            r = F.conv2d(r, sd[p + "5.weight"], sd.get(p + "5.bias"))
            t = tap(t + inorm(r))
        m += n_blocks
        t = tap(blur_up(t)); t = tap(conv(t, m + 1, 1)); t = tap(inorm(t)); t = tap(F.relu(t))
        t = tap(blur_up(t)); t = tap(conv(t, m + 5, 1)); t = tap(inorm(t)); t = tap(F.relu(t))
        t = tap(F.pad(t, (3, 3, 3, 3), mode="reflect"))
        t = tap(conv(t, m + 9))
        t = tap(torch.tanh(t))
    except _Done:
        return feats
    if len(layers):
        return t, feats
    return t


def unet_custom_forward(sd, x, num_downs=8, style_code=None, num_layer_style_code=1):
    """CustomUnetGenerator.forward (models/networks.py:1576-1645) with Down/Up blocks
    (thirdparty/unet/unet_parts_custom.py:9-80).  Mirrors the in-place LeakyReLU aliasing of
    the skip tensors: every skip is the LeakyReLU'd activation.
    style_code (skitG, 'concat'+'tile' mode, networks.py:1600-1623) is tiled and concatenated
    to the decoder input of the innermost `num_layer_style_code` levels."""
    skips = []
    t = x
    for i in range(num_downs):
        if i > 0:
            t = F.leaky_relu(t, 0.2)
            skips[-1] = t  # in-place activation aliases the stored skip
        key = "down%d.model.%d" % (i, 0 if i == 0 else 1)
        t = F.conv2d(t, sd[key + ".weight"], sd[key + ".bias"], stride=2, padding=1)
        if 0 < i < num_downs - 1:
            t = inorm(t)
        skips.append(t)

    def up(name, t_in, skip, i):
        if not (i == 0 or i == num_downs - 1):
            t_in = torch.cat([t_in, skip], dim=1)
        t_in = F.relu(t_in)
        y = F.conv_transpose2d(t_in, sd[name + ".model.1.weight"], sd[name + ".model.1.bias"], stride=2, padding=1)
        return torch.tanh(y) if i == 0 else inorm(y)

    t_T = None
    for i in range(num_downs - 1, -1, -1):
        if style_code is not None and i >= num_downs - num_layer_style_code:
            sc = style_code.to(torch.float32)[:, :, None, None].expand(-1, -1, t.shape[2], t.shape[3])
            t = torch.cat([t, sc], dim=1)
            if t_T is not None:
                t_T = torch.cat([t_T, sc], dim=1)
        skip = skips[i]
        if ("up%d_T.model.1.weight" % i) in sd:
            if t_T is None:
                t_T = t
            t_T = up("up%d_T" % i, t_T, skip, i)
        t = up("up%d" % i, t, skip, i)
    if t_T is not None:
        t = torch.cat([t, t_T], dim=1)
    return t


# --------------------------------------------------------------------------- discriminators
# --------------------------------------------------------------------------- StyleGAN2 generator (a4)
def sg2_channels(ngf):
    """Channel table of StyleGAN2Encoder / StyleGAN2Decoder (models/stylegan_networks.py:805-816, 862-873)."""
    m = ngf / 32
    t = {4: min(512, int(round(4096 * m))), 8: min(512, int(round(2048 * m))), 16: min(512, int(round(1024 * m))),
         32: min(512, int(round(512 * m)))}
    t.update({64: int(round(256 * m)), 128: int(round(128 * m)), 256: int(round(64 * m)), 512: int(round(32 * m)),
              1024: int(round(16 * m))})
    return t


def sg2_blur(x, pad0, pad1, gain=1.0):
    """Blur([1,3,3,1], pad) = upfirdn2d(up=1, down=1) (stylegan_networks.py:38-72, 140-157): zero pad by (pad0, pad1) on
    both axes, then a VALID 4x4 correlation with the flipped (symmetric) normalised kernel, per channel."""
    k1 = torch.tensor([1.0, 3.0, 3.0, 1.0])
    k = (k1[None, :] * k1[:, None])
    k = k / k.sum() * gain
    c = x.shape[1]
    xp = F.pad(x, [pad0, pad1, pad0, pad1])
    return F.conv2d(xp, torch.flip(k, [0, 1])[None, None].repeat(c, 1, 1, 1), groups=c)


def sg2_flrelu(x, bias):
    """FusedLeakyReLU (stylegan_networks.py:18-35): leaky_relu(x + b, 0.2) * sqrt(2)."""
    return F.leaky_relu(x + bias.view(1, -1, 1, 1), 0.2) * (2 ** 0.5)


def sg2_conv_layer(sd, pre, x, k, downsample=False, activate=True):
    """ConvLayer (stylegan_networks.py:613-665): [Blur] + EqualConv2d (weight * 1/sqrt(ci*k*k), :179-190) [+ FusedLeakyReLU].
    Sequential indices: blur 0 / conv 1 / act 2 when downsampling, conv 0 / act 1 otherwise; every ConvLayer of the
    generator has bias=True, so the conv itself carries a bias only when there is no activation — never here (skip convs
    are built with bias=False)."""
    i = 0
    if downsample:
        p = (4 - 2) + (k - 1)
        x = sg2_blur(x, (p + 1) // 2, p // 2)
        i = 1
    w = sd[pre + "%d.weight" % i]
    scale = 1.0 / math.sqrt(w.shape[1] * k * k)
    x = F.conv2d(x, w * scale, None, stride=2 if downsample else 1, padding=0 if downsample else k // 2)
    if activate:
        x = sg2_flrelu(x, sd[pre + "%d.bias" % (i + 1)])
    return x


def sg2_resblock(sd, pre, x, downsample):
    """ResBlock (stylegan_networks.py:668-689), skip_gain 1: (conv2(conv1(x)) + skip(x)) / sqrt(2); the skip is a 1x1
    ConvLayer (blur + stride 2, no bias, no activation) when the block downsamples or changes width, identity otherwise."""
    out = sg2_conv_layer(sd, pre + "conv1.", x, 3)
    out = sg2_conv_layer(sd, pre + "conv2.", out, 3, downsample=downsample)
    skip = x
    if (pre + "skip.1.weight") in sd or (pre + "skip.0.weight") in sd:
        skip = sg2_conv_layer(sd, pre + "skip.", x, 1, downsample=downsample, activate=False)
    return (out + skip) / math.sqrt(2.0)


def sg2_styled_conv_up(sd, pre, x, noise=None):
    """StyledConv(upsample=True, style=None) (stylegan_networks.py:375-412) around ModulatedConv2d (:248-348): style = ones,
    weight = W / sqrt(ci*9), demodulated per output channel (rsqrt(sum w^2 + 1e-8)); conv_transpose2d stride 2, no padding;
    Blur(pad=(1, 1)) with the kernel scaled by 4; + noise_weight * noise; FusedLeakyReLU."""
    w = sd[pre + "conv.weight"][0]                      # [co, ci, 3, 3]
    co, ci, k, _ = w.shape
    w = w * (1.0 / math.sqrt(ci * k * k))
    w = w * torch.rsqrt(w.pow(2).sum([1, 2, 3]) + 1e-8).view(co, 1, 1, 1)
    out = F.conv_transpose2d(x, w.transpose(0, 1), padding=0, stride=2)
    p = (4 - 2) - (k - 1)
    out = sg2_blur(out, (p + 1) // 2 + 1, p // 2 + 1, gain=4.0)
    if noise is not None:
        out = out + sd[pre + "noise.weight"] * noise
    return sg2_flrelu(out, sd[pre + "activate.bias"])


def stylegan2_g_forward(sd, x, n_blocks=6, num_downsampling=1, layers=(), encode_only=False, noises=None):
    """StyleGAN2Generator.forward (models/stylegan_networks.py:912-929; encoder :800-851, decoder :854-909).
    sd: reference state_dict.  noises: one [n,1,H,W] tensor per StyledConv, or None for the noise-free ('small') variant /
    zero noise weights — the reference draws them from the global RNG inside NoiseInjection (:351-362).
    Returns fake [n,3,S,S] (and the encoder features at `layers`, indices into encoder.convs)."""
    layers = list(layers)
    feat, feats = x, []
    n_enc = 2 + num_downsampling + n_blocks // 2
    if -1 in layers:
        layers.append(n_enc - 1)
    for li in range(n_enc):
        pre = "encoder.convs.%d." % li
        if li == 1:
            feat = sg2_conv_layer(sd, pre, feat, 1)
        elif li >= 2:
            feat = sg2_resblock(sd, pre, feat, downsample=li < 2 + num_downsampling)
        if li in layers:
            feats.append(feat)
    if encode_only:
        return feats
    for li in range(n_blocks // 2):
        feat = sg2_resblock(sd, "decoder.convs.%d." % li, feat, downsample=False)
    for j in range(num_downsampling):
        li = n_blocks // 2 + j
        feat = sg2_styled_conv_up(sd, "decoder.convs.%d." % li, feat, None if noises is None else noises[j])
    feat = sg2_conv_layer(sd, "decoder.convs.%d." % (n_blocks // 2 + num_downsampling), feat, 1)
    return (feat, feats) if len(layers) > 0 else feat


# --------------------------------------------------------------------------- LPIPS (third-party: pip lpips 0.1.4)
# The reference builds `lpips.LPIPS(net="vgg")` (models/sinskitG_model.py:495) and calls it on (fake_I, real_I) (:1711) and on
# single-channel 32x32 touch patches (:1639-1645).  The package is NOT vendored under /root/reference and not installed in
# this image; the algorithm below restates its published code (lpips/lpips.py: ScalingLayer, normalize_tensor, NetLinLayer,
# spatial_average; lpips/pretrained_networks.py: vgg16 slices = torchvision vgg16.features[0:4], [4:9], [9:16], [16:23],
# [23:30]).  Pinning: the VGG16 trunk is checked against torchvision's own vgg16 module (tests/test_oracle_golden.py);
# the LPIPS head has no fixture to check against -> "parity unpinned" for the head.  Weights: the pretrained VGG16 / lin
# checkpoints are not available offline; tests and bench use random weights with the package's state_dict keys.
LPIPS_SHIFT = (-0.030, -0.088, -0.188)
LPIPS_SCALE = (0.458, 0.448, 0.450)
VGG_SLICES = ((0, 2), (5, 7), (10, 12, 14), (17, 19, 21), (24, 26, 28))   # conv indices in vgg16.features per slice
VGG_CHNS = (64, 128, 256, 512, 512)


def vgg16_features(sd, x, prefix="net."):
    """relu1_2, relu2_2, relu3_3, relu4_3, relu5_3 of a VGG16 whose convs live at `net.slice{s}.{idx}.{weight,bias}`."""
    outs = []
    h = x
    for s, idxs in enumerate(VGG_SLICES):
        if s > 0:
            h = F.max_pool2d(h, 2, 2)
        for i in idxs:
            h = F.relu(F.conv2d(h, sd["%sslice%d.%d.weight" % (prefix, s + 1, i)], sd["%sslice%d.%d.bias" % (prefix, s + 1, i)], padding=1))
        outs.append(h)
    return outs


def lpips_vgg(sd, in0, in1):
    """LPIPS(net='vgg', lpips=True, spatial=False).forward(in0, in1, normalize=False) -> [N,1,1,1]."""
    shift = torch.tensor(LPIPS_SHIFT).view(1, 3, 1, 1)
    scale = torch.tensor(LPIPS_SCALE).view(1, 3, 1, 1)
    f0 = vgg16_features(sd, (in0 - shift) / scale)      # a 1-channel input broadcasts to 3 channels here
    f1 = vgg16_features(sd, (in1 - shift) / scale)
    val = 0
    for k in range(5):
        n0 = f0[k] / (torch.sqrt(torch.sum(f0[k] ** 2, dim=1, keepdim=True)) + 1e-10)
        n1 = f1[k] / (torch.sqrt(torch.sum(f1[k] ** 2, dim=1, keepdim=True)) + 1e-10)
        d = (n0 - n1) ** 2
        val = val + F.conv2d(d, sd["lin%d.model.1.weight" % k]).mean([2, 3], keepdim=True)
    return val


def lpips_random_state(seed=0, weight_gain=1.0):
    """Random LPIPS-VGG weights with the package's state_dict keys (no pretrained checkpoint offline): He-normal convs,
    small biases, non-negative lin weights (the trained ones are clamped to >= 0)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    cin = 3
    for s, idxs in enumerate(VGG_SLICES):
        for i in idxs:
            co = VGG_CHNS[s]
            sd["net.slice%d.%d.weight" % (s + 1, i)] = torch.randn(co, cin, 3, 3, generator=g) * math.sqrt(2.0 / (cin * 9)) * weight_gain
            sd["net.slice%d.%d.bias" % (s + 1, i)] = torch.randn(co, generator=g) * 0.05
            cin = co
        sd["lin%d.model.1.weight" % s] = torch.rand(1, VGG_CHNS[s], 1, 1, generator=g) / VGG_CHNS[s] * 4
    return sd


def nlayer_d_forward(sd, prefix, x, n_layers=3, bn_momentum=0.1, update_running=True):
    """NLayerDiscriminator.forward, BatchNorm2d in training mode
    (models/networks.py:1702-1750; index map SURVEY.md A.5).  Running stats in `sd` are
    updated in place like nn.BatchNorm2d does (momentum 0.1, unbiased running var)."""
    def bn(t, i):
        p = "%s%d." % (prefix, i)
        mu = t.mean(dim=(0, 2, 3))
        var = t.var(dim=(0, 2, 3), unbiased=False)
        if update_running and (p + "running_mean") in sd:
            n = t.numel() / t.shape[1]
            with torch.no_grad():
                sd[p + "running_mean"].mul_(1 - bn_momentum).add_(bn_momentum * mu.detach())
                sd[p + "running_var"].mul_(1 - bn_momentum).add_(bn_momentum * var.detach() * n / max(n - 1, 1))
                sd[p + "num_batches_tracked"] += 1
        th = (t - mu[None, :, None, None]) / torch.sqrt(var[None, :, None, None] + EPS_NORM)
        return th * sd[p + "weight"][None, :, None, None] + sd[p + "bias"][None, :, None, None]

    def conv(t, i, s):
        return F.conv2d(t, sd["%s%d.weight" % (prefix, i)], sd["%s%d.bias" % (prefix, i)], stride=s, padding=2)

    t = F.leaky_relu(conv(x, 0, 2), 0.2)
    i = 2
    for _ in range(1, n_layers):
        t = F.leaky_relu(bn(conv(t, i, 2), i + 1), 0.2)
        i += 3
    t = F.leaky_relu(bn(conv(t, i, 1), i + 1), 0.2)
    i += 3
    return conv(t, i, 1)


def multiscale_d_forward(sd, x, num_D=3, n_layers=3, update_running=True):
    """MultiscaleDiscriminator.forward (models/networks.py:1681-1693): scale i runs
    `layer{num_D-1-i}` on the i-times AvgPool2d(3,2,1,count_include_pad=False) input."""
    out = []
    t = x
    for i in range(num_D):
        out.append([nlayer_d_forward(sd, "layer%d." % (num_D - 1 - i), t, n_layers, update_running=update_running)])
        if i != num_D - 1:
            t = F.avg_pool2d(t, 3, stride=2, padding=1, count_include_pad=False)
    return out


def gan_loss(pred, target_is_real, gan_mode="nonsaturating", real_label=1.0, fake_label=0.0):
    """GANLoss.__call__ (models/networks.py:500-542).  Multiscale (list of lists) sums the
    per-sample losses of the last map of each scale; a bare tensor uses `input[-1]`, i.e.
    the LAST BATCH ELEMENT only (reference quirk, :541-542)."""
    def single(p):
        bs = p.shape[0]
        if gan_mode == "nonsaturating":
            return F.softplus(-p if target_is_real else p).view(bs, -1).mean(dim=1)
        if gan_mode == "hinge":
            return F.relu(1.0 - p if target_is_real else 1.0 + p).view(bs, -1).mean(dim=1)
        if gan_mode in ("wgan", "wgangp"):
            return -p.mean() if target_is_real else p.mean()
        t = torch.full_like(p, real_label if target_is_real else fake_label)        # get_target_tensor (:482-497)
        if gan_mode == "lsgan":
            return F.mse_loss(p, t)
        if gan_mode == "vanilla":
            return F.binary_cross_entropy_with_logits(p, t)
        raise NotImplementedError(gan_mode)

    if isinstance(pred[0], list):
        total = 0
        for scale in pred:
            total = total + single(scale[-1])
        return total
    return single(pred[-1])


# --------------------------------------------------------------------------- patches
def patch_offsets_from_coords(coords, scale_multiplier=1):
    """find_coords_for_patch (models/model_utils.py:37-57): float64 numpy arithmetic,
    np.round (half-to-even), then float32 -> int32."""
    c = np.squeeze(np.asarray(coords, dtype=np.float64))
    if c.ndim == 1:
        c = c[None]
    ox = np.round((c[..., 0] + c[..., -2] / c[..., -3]) * scale_multiplier)
    oy = np.round((c[..., 1] + c[..., -1] / c[..., -3]) * scale_multiplier)
    cs = np.round(c[..., -4] / c[..., -3] * scale_multiplier)
    to_i = lambda a: np.asarray(a, dtype=np.float32).astype(np.int32)
    return to_i(ox), to_i(oy), to_i(cs)


def gather_patches(img, ox, oy, cutout):
    """The gather of get_patch_in_input (models/model_utils.py:254-335): patch p, pixel (y,x) =
    img[0, :, clamp(y+oy[p],0,H-1), clamp(x+ox[p],0,W-1)] for y,x in [0, max(cutout))."""
    assert img.shape[0] == 1
    H, W = img.shape[-2:]
    ps = int(np.max(cutout))
    ar = torch.arange(ps, device=img.device)
    ys = (torch.as_tensor(np.asarray(oy), dtype=torch.long, device=img.device)[:, None] + ar[None]).clamp(0, H - 1)  # [P, ps]
    xs = (torch.as_tensor(np.asarray(ox), dtype=torch.long, device=img.device)[:, None] + ar[None]).clamp(0, W - 1)
    out = img[0][:, ys[:, :, None], xs[:, None, :]]  # [C, P, ps, ps]
    return out.permute(1, 0, 2, 3).contiguous()


def random_patch_offsets(M, sample_size, rng=_pyrandom):
    """Random-mode offsets of get_patch_in_input (models/model_utils.py:212-222): a 17x17
    all-ones conv with padding 1 over M (map shrinks to (H-14)x(W-14)), clamp to [0,1],
    torch.nonzero rows in row-major order, `random.sample(range(nnz), sample_size)`; the
    (row, col) of the shrunken map are used directly as (offset_y, offset_x)."""
    k = torch.ones(1, 1, 17, 17, dtype=M.dtype)
    er = torch.clamp(F.conv2d(M[:, :1], k, padding=1), 0, 1)
    nz = torch.nonzero(er, as_tuple=False)
    pick = rng.sample(range(nz.shape[0]), sample_size)
    sel = nz[pick][:, -2:]
    return sel[:, 1].to(torch.int32).numpy(), sel[:, 0].to(torch.int32).numpy()


def get_patch_in_input(img, coords=None, sample_size=None, scale_multiplier=1, patch_size=32,
                       offset_x=None, offset_y=None, M=None, return_offset=False, rng=_pyrandom):
    """get_patch_in_input (models/model_utils.py:72-405) for the three usages on the hot path:
    known coords; random offsets inside the mask; caller-supplied offsets."""
    patch_size = patch_size * scale_multiplier
    if coords is not None:
        assert np.asarray(coords).shape[0] == 1, "coords should have batch size of 1"
        ox, oy, cs = patch_offsets_from_coords(coords, scale_multiplier)
    else:
        assert sample_size is not None
        if offset_x is None:
            ox, oy = random_patch_offsets(M, sample_size, rng)
        else:
            ox, oy = np.asarray(offset_x).reshape(-1), np.asarray(offset_y).reshape(-1)
        cs = np.full((sample_size,), patch_size, dtype=np.int32)
    out = gather_patches(img, ox, oy, cs)
    if out.shape[-1] < patch_size:
        out = F.interpolate(out, size=(patch_size, patch_size), mode="bicubic", align_corners=False, antialias=True)
    if return_offset:
        return out, ox / scale_multiplier, oy / scale_multiplier, cs / scale_multiplier
    return out


def compute_normal(T, scale_nz=0.25):
    """compute_normal (models/model_utils.py:418-425): F.normalize([gx, gy, scale_nz], dim=1)."""
    n = torch.cat([T[:, 0:1], T[:, 1:2], torch.full_like(T[:, 0:1], scale_nz)], dim=1)
    return n / n.norm(dim=1, keepdim=True).clamp_min(1e-12)


def diffaugment_bs(x, u_b, u_s):
    """DiffAugment policy 'bs' (thirdparty/DiffAugment.py:25-33): brightness x + (U_b - 0.5),
    then saturation (x - mean_c) * 2 U_s + mean_c.  U_* are the per-sample torch.rand draws."""
    u_b = torch.as_tensor(u_b, dtype=x.dtype, device=x.device).view(-1, 1, 1, 1)
    u_s = torch.as_tensor(u_s, dtype=x.dtype, device=x.device).view(-1, 1, 1, 1)
    x = x + (u_b - 0.5)
    m = x.mean(dim=1, keepdim=True)
    return ((x - m) * (u_s * 2) + m).contiguous()


# --------------------------------------------------------------------------- PatchNCE
def patch_sample_f(feats, patch_ids, mlps=None):
    """PatchSampleF.forward with given ids (models/networks.py:689-719) + Normalize (:585-594).
    mlps: optional list of (W1, b1, W2, b2) per feature (Linear-ReLU-Linear)."""
    outs = []
    for i, f in enumerate(feats):
        B, C, H, W = f.shape
        fr = f.permute(0, 2, 3, 1).flatten(1, 2)
        ids = torch.as_tensor(np.asarray(patch_ids[i]), dtype=torch.long, device=f.device)
        xs = fr[:, ids, :].flatten(0, 1)
        if mlps is not None:
            W1, b1, W2, b2 = mlps[i]
            xs = F.linear(F.relu(F.linear(xs, W1, b1)), W2, b2)
        nrm = xs.pow(2).sum(1, keepdim=True).pow(0.5)
        outs.append(xs / (nrm + 1e-7))
    return outs


def patchnce_loss(feat_q, feat_k, nce_T=0.07, batch_size=1, all_negatives_from_minibatch=False):
    """PatchNCELoss.forward (models/patchnce.py:13-55)."""
    P, dim = feat_q.shape
    feat_k = feat_k.detach()
    l_pos = (feat_q * feat_k).sum(dim=1, keepdim=True)
    b = 1 if all_negatives_from_minibatch else batch_size
    q = feat_q.view(b, -1, dim)
    k = feat_k.view(b, -1, dim)
    n = q.shape[1]
    l_neg = torch.bmm(q, k.transpose(1, 2))
    l_neg = l_neg.masked_fill(torch.eye(n, dtype=torch.bool, device=q.device)[None], -10.0).view(-1, n)
    logits = torch.cat([l_pos, l_neg], dim=1) / nce_T
    return torch.logsumexp(logits, dim=1) - logits[:, 0]


# --------------------------------------------------------------------------- optimiser
def adam_step(p, g, m, v, step, lr, beta1=0.0, beta2=0.99, eps=1e-8):
    """torch.optim.Adam (no weight decay, no amsgrad) as configured at
    models/sinskitG_model.py:590-599: returns nothing, updates p/m/v in place."""
    m.mul_(beta1).add_(g, alpha=1 - beta1)
    v.mul_(beta2).addcmul_(g, g, value=1 - beta2)
    bc1 = 1 - beta1 ** step
    bc2 = 1 - beta2 ** step
    denom = (v.sqrt() / math.sqrt(bc2)).add_(eps)
    p.addcdiv_(m, denom, value=-lr / bc1)


def linear_lr_factor(epoch, epoch_count=1, n_epochs=5, n_epochs_decay=400):
    """get_scheduler('linear') lambda (models/networks.py:161-165)."""
    return 1.0 - max(0, epoch + epoch_count - n_epochs) / float(n_epochs_decay + 1)


# --------------------------------------------------------------------------- the train step
class StepConfig:
    """The sinskitG options the step reads (defaults: models/sinskitG_model.py:50-357)."""

    def __init__(self, **kw):
        self.netG = "resnet_9blocks"
        self.gan_mode = "nonsaturating"      # 'hinge' is the only other mode the reference's own G2 loss survives (len() of a 0-d tensor, :1783)
        self.n_blocks = 9
        self.num_D = 3
        self.n_layers_D = 3
        self.lambda_G1_GAN = 1.0
        self.lambda_G1_L1 = 100.0
        self.lambda_G2_GAN = 5.0
        self.lambda_G2_L1 = 10.0
        self.batch_size_G2 = 64
        self.add_fake_T_sample_size = 32
        self.scale_nz = 0.25
        self.lr = 1e-3
        self.lr_G2 = 5e-4
        self.beta1 = 0.0
        self.beta2 = 0.99
        self.use_diffaug = True
        self.use_more_fakeT = True
        # PatchNCE wiring (new behaviour): CUT-style, off by default
        # LPIPS-VGG terms (reference defaults 1.0 / 10.0, sinskitG_model.py:69-72, 105-108); off unless sdL is given
        self.lambda_G1_lpips = 0.0
        self.lambda_G2_lpips = 0.0
        self.lambda_NCE = 0.0
        self.nce_layers = (0, 4, 8, 12, 16)
        self.nce_T = 0.07
        self.num_patches = 256
        self.__dict__.update(kw)


def g_forward(cfg, sdG, x, **kw):
    if cfg.netG.startswith("resnet"):
        return resnet_g_forward(sdG, x, n_blocks=cfg.n_blocks, **kw)
    if cfg.netG == "unet256_custom":
        return unet_custom_forward(sdG, x, **kw)
    raise NotImplementedError("Generator model name [%s] is not recognized" % cfg.netG)


def model_forward(cfg, sdG, real_S, S_pe, M, real_I=None, rand=None, style_code=None):
    """SinSKITGModel.forward (models/sinskitG_model.py:1293-1344).  `real_S`/`real_I` are
    already mask-multiplied by set_input (:724,734).  rand = dict(real_b, real_s, fake_b, fake_s)."""
    kw = {}
    if style_code is not None and cfg.netG == "unet256_custom":   # skitG: the style code enters the U-Net generators only
        kw = dict(style_code=style_code, num_layer_style_code=getattr(cfg, "num_layer_style_code", 1))
    out = g_forward(cfg, sdG, torch.cat([real_S, S_pe], 1), **kw)
    fake_I = out[:, 0:3] * M
    fake_T = out[:, -2:] * M
    res = dict(fake_I=fake_I, fake_T=fake_T, fake_N=compute_normal(fake_T.detach(), cfg.scale_nz))
    if real_I is not None:
        if cfg.use_diffaug:
            res["aug_real_I"] = diffaugment_bs(real_I, rand["real_b"], rand["real_s"]) * M
            res["aug_fake_I"] = diffaugment_bs(fake_I, rand["fake_b"], rand["fake_s"]) * M
        else:
            res["aug_real_I"] = real_I * M
            res["aug_fake_I"] = fake_I * M
    return res


def train_step(cfg, sdG, sdD, sdD2, opt_state, batch, rand, step=1, lr_factor=1.0, grad_hook=None, sdF=None, sdL=None):
    """SinSKITGModel.optimize_parameters (models/sinskitG_model.py:601-700) with
    compute_D1_loss (:1346-1407), compute_D2_loss (:1409-1617), compute_G1_loss (:1660-1726),
    compute_G2_loss (:1728-1842); LPIPS / vision-aided terms off (SURVEY.md §8c flags).

    batch: real_S [1,1,S,S] (pre-mask), real_I [1,3,S,S], M [1,1,S,S], T_images [NT,2,32,32],
           I_masks [NT,1,32,32], T_coords [1,NT,8] (float64 numpy).
    rand:  real_b, real_s, fake_b, fake_s (DiffAugment draws), fake_ox, fake_oy (random patch
           offsets for the NF extra fake patches).
    grad_hook(name, dict_of_grads): optional, e.g. gradient all-reduce in the DDP tests.
    Mutates sdG/sdD/sdD2/opt_state in place; returns dict of losses and tensors."""
    M = batch["M"]
    real_S = batch["real_S"] * M
    real_I = batch["real_I"] * M
    real_T = batch["T_images"] * batch["I_masks"]
    I_masks = batch["I_masks"]
    coords = batch["T_coords"]
    S_pe = spe_grid(real_S.shape[2], real_S.shape[3], 4, real_S.shape[0]).to(real_S.device)
    NT = cfg.batch_size_G2

    gp = {k: v.requires_grad_(True) for k, v in sdG.items() if v.dtype.is_floating_point and "filt" not in k and "style_code_mapping" not in k}
    sdG_run = dict(sdG)
    sdG_run.update(gp)
    fw = model_forward(cfg, sdG_run, real_S, S_pe, M, real_I, rand, style_code=batch.get("style_code"))
    fake_I, fake_T = fw["fake_I"], fw["fake_T"]

    ox, oy, cs = patch_offsets_from_coords(coords)
    fake_T_p = gather_patches(fake_T, ox, oy, cs)
    S_p = gather_patches(real_S, ox, oy, cs).detach()
    areal_p = torch.cat([gather_patches(fw["aug_real_I"], ox, oy, cs).detach(), I_masks], 1)
    afake_p = torch.cat([gather_patches(fw["aug_fake_I"], ox, oy, cs).detach(), I_masks], 1)
    losses = {}

    def run_opt(name, sd, lr, loss, grads=None):
        params = {k: v for k, v in sd.items() if v.requires_grad}
        if grads is None:
            grads = torch.autograd.grad(loss, list(params.values()), allow_unused=True)
        gd = {k: g for k, g in zip(params.keys(), grads) if g is not None}
        if grad_hook is not None:
            grad_hook(name, gd)
        st = opt_state.setdefault(name, {})
        with torch.no_grad():
            for k in gd:
                if k not in st:
                    st[k] = (torch.zeros_like(sd[k]), torch.zeros_like(sd[k]))
            if getattr(cfg, "foreach_adam", False) and gd:
                # the same update through torch's multi-tensor primitives, as torch.optim.Adam(foreach=True) — the default on
                # CUDA — issues it (bench.py's eager-PyTorch-on-B200 arm; one launch per op instead of one per tensor)
                ks = list(gd)
                ps, gs = [sd[k] for k in ks], [gd[k] for k in ks]
                ms, vs = [st[k][0] for k in ks], [st[k][1] for k in ks]
                torch._foreach_mul_(ms, cfg.beta1)
                torch._foreach_add_(ms, gs, alpha=1 - cfg.beta1)
                torch._foreach_mul_(vs, cfg.beta2)
                torch._foreach_addcmul_(vs, gs, gs, value=1 - cfg.beta2)
                den = torch._foreach_sqrt(vs)
                torch._foreach_div_(den, math.sqrt(1 - cfg.beta2 ** step))
                torch._foreach_add_(den, 1e-8)
                torch._foreach_addcdiv_(ps, ms, den, value=-(lr * lr_factor) / (1 - cfg.beta1 ** step))
            else:
                for k, g in gd.items():
                    adam_step(sd[k], g, st[k][0], st[k][1], step, lr * lr_factor, cfg.beta1, cfg.beta2)
        for v in sd.values():
            if v.dtype.is_floating_point:
                v.requires_grad_(False)
        return gd

    def d_params_on(sd):
        for k, v in sd.items():
            if v.dtype.is_floating_point and "running" not in k:
                v.requires_grad_(True)

    # ---- D1 step
    d_params_on(sdD)
    pf = multiscale_d_forward(sdD, torch.cat([real_S, fake_I.detach()], 1), cfg.num_D, cfg.n_layers_D)
    l_f = gan_loss(pf, False, cfg.gan_mode).mean() * cfg.lambda_G1_GAN
    pr = multiscale_d_forward(sdD, torch.cat([real_S, real_I], 1), cfg.num_D, cfg.n_layers_D)
    l_r = gan_loss(pr, True, cfg.gan_mode).mean() * cfg.lambda_G1_GAN
    losses["D_fake_I"], losses["D_real_I"] = l_f.item(), l_r.item()
    grads_D = run_opt("D", sdD, cfg.lr, (l_f + l_r) * 0.5)

    # ---- D2 step
    d_params_on(sdD2)
    fake_in = torch.cat([fake_T_p.detach(), S_p, afake_p], 1)
    l_f2 = gan_loss(multiscale_d_forward(sdD2, fake_in, cfg.num_D, cfg.n_layers_D), False, cfg.gan_mode).mean() * cfg.lambda_G2_GAN
    # full-resolution D2 pass (:1495): visualisation only, but it updates BN running stats
    full_in = torch.cat([fake_T.detach(), real_S, fw["aug_fake_I"].detach(), M], 1)
    with torch.no_grad():
        pred_full = multiscale_d_forward(sdD2, full_in, cfg.num_D, cfg.n_layers_D)[-1][-1]
    l_m2 = 0.0
    if cfg.use_more_fakeT:
        fox, foy = rand["fake_ox"], rand["fake_oy"]
        NF = cfg.add_fake_T_sample_size
        csf = np.full((NF,), 32, dtype=np.int32)
        more = torch.cat([gather_patches(fake_T.detach(), fox, foy, csf),
                          gather_patches(real_S, fox, foy, csf),
                          gather_patches(fake_I.detach(), fox, foy, csf),
                          torch.ones(NF, 1, 32, 32, device=real_S.device)], 1)
        l_m2 = gan_loss(multiscale_d_forward(sdD2, more, cfg.num_D, cfg.n_layers_D), False, cfg.gan_mode).mean() * cfg.lambda_G2_GAN
        losses["D_more_fake_T"] = l_m2.item()
    real_in = torch.cat([real_T, S_p, areal_p], 1)
    l_r2 = gan_loss(multiscale_d_forward(sdD2, real_in, cfg.num_D, cfg.n_layers_D), True, cfg.gan_mode).mean() * cfg.lambda_G2_GAN
    losses["D_fake_T_concat"], losses["D_real_T_concat"] = l_f2.item(), l_r2.item()
    grads_D2 = run_opt("D2", sdD2, cfg.lr_G2, (l_f2 + l_m2 + l_r2) * 0.5)

    # ---- G step (D, D2 already updated; their params frozen)
    pg = multiscale_d_forward(sdD, torch.cat([real_S, fake_I], 1), cfg.num_D, cfg.n_layers_D)
    l_gan = gan_loss(pg, True, cfg.gan_mode).mean() * cfg.lambda_G1_GAN
    l_l1 = (fake_I - real_I).abs().mean() * cfg.lambda_G1_L1
    with torch.no_grad():  # value only: computed on a detached clone (:1751,1781)
        pg2 = multiscale_d_forward(sdD2, fake_in, cfg.num_D, cfg.n_layers_D)
        l_g2 = (gan_loss(pg2, True, cfg.gan_mode) * cfg.lambda_G2_GAN).view(-1, NT).mean(dim=0).sum()
    l_l1_2 = ((fake_T_p - real_T).abs() * cfg.lambda_G2_L1).view(-1, NT, *fake_T_p.shape[1:]).sum(dim=1).mean()
    losses.update(G_GAN=l_gan.item(), G_L1=l_l1.item(), G2_GAN=l_g2.item(), G2_L1=l_l1_2.item())
    loss_G = l_gan + l_l1 + l_g2 + l_l1_2
    if sdL is not None and cfg.lambda_G1_lpips > 0:      # compute_G1_loss :1707-1715
        l_lp = lpips_vgg(sdL, fake_I, real_I).mean() * cfg.lambda_G1_lpips
        losses["G_lpips"] = l_lp.item()
        loss_G = loss_G + l_lp
    if sdL is not None and cfg.lambda_G2_lpips > 0:      # _compute_touch_lpips_loss :1619-1658: gx and gy separately, 1 channel each
        l_gx = lpips_vgg(sdL, fake_T_p[:, 0:1], real_T[:, 0:1]).view(-1, NT, 1, 1, 1).sum(dim=1).mean()
        l_gy = lpips_vgg(sdL, fake_T_p[:, 1:2], real_T[:, 1:2]).view(-1, NT, 1, 1, 1).sum(dim=1).mean()
        l_lp2 = cfg.lambda_G2_lpips * (l_gx + l_gy)
        losses["G2_lpips"] = l_lp2.item()
        loss_G = loss_G + l_lp2
    if cfg.lambda_NCE > 0:
        # CUT-style PatchNCE wiring (NEW behaviour: PatchNCELoss / PatchSampleF are dead code in the reference, SURVEY.md 0.4;
        # the functions themselves are pinned against the reference in tests/golden/ops.npz).  keys: generator features of
        # the input; query: the same encoder on the channel mean of the generated image (+ the same positional encoding).
        layers = sorted(cfg.nce_layers)
        with torch.no_grad():
            feat_k = resnet_g_forward(sdG_run, torch.cat([real_S, S_pe], 1), cfg.n_blocks, layers=layers, encode_only=True)
        S_q = fake_I.mean(1, keepdim=True)
        feat_q = resnet_g_forward(sdG_run, torch.cat([S_q, S_pe], 1), cfg.n_blocks, layers=layers, encode_only=True)
        mlps = None
        if sdF is not None:   # netF = 'mlp_sample': per-layer Linear-ReLU-Linear (networks.py:678-686), trained by its own Adam
            for v in sdF.values():
                v.requires_grad_(True)
            mlps = [(sdF["mlp_%d.0.weight" % i], sdF["mlp_%d.0.bias" % i], sdF["mlp_%d.2.weight" % i], sdF["mlp_%d.2.bias" % i])
                    for i in range(len(layers))]
        kp = patch_sample_f(feat_k, rand["nce_ids"], mlps)
        qp = patch_sample_f(feat_q, rand["nce_ids"], mlps)
        l_nce = sum(patchnce_loss(q, k, cfg.nce_T, batch_size=real_S.shape[0]).mean() for q, k in zip(qp, kp)) / len(layers) * cfg.lambda_NCE
        losses["NCE"] = l_nce.item()
        loss_G = loss_G + l_nce
    for k in gp:
        sdG[k] = gp[k]
    grads_F = None
    use_F = cfg.lambda_NCE > 0 and sdF is not None
    if use_F:   # both gradient sets come from the same graph: take F's first, apply its Adam after G's
        f_raw = torch.autograd.grad(loss_G, [v for v in sdF.values() if v.requires_grad], allow_unused=True, retain_graph=True)
    grads_G = run_opt("G", sdG, cfg.lr, loss_G)
    if use_F:
        grads_F = run_opt("F", sdF, cfg.lr, None, grads=f_raw)

    return dict(losses=losses, fake_I=fake_I.detach(), fake_T=fake_T.detach(), fake_N=fw["fake_N"],
                aug_fake_I=fw["aug_fake_I"].detach(), aug_real_I=fw["aug_real_I"].detach(),
                fake_T_patches=fake_T_p.detach(), pred_fake_T_full=pred_full,
                grads_G=grads_G, grads_D=grads_D, grads_D2=grads_D2, grads_F=grads_F)


# --------------------------------------------------------------------------- synthetic data
def synthetic_batch(S, NT=64, seed=0, ellipse_mask=False):
    """The seeded synthetic batch of SURVEY.md §8(d): the dict `set_input` consumes
    (models/sinskitG_model.py:702-793), single image."""
    g = torch.Generator().manual_seed(seed)
    rs = np.random.RandomState(seed)
    M = torch.ones(1, 1, S, S)
    if ellipse_mask:
        yy, xx = torch.meshgrid(torch.arange(S), torch.arange(S), indexing="ij")
        M = ((((yy - S / 2) / (0.45 * S)) ** 2 + ((xx - S / 2) / (0.40 * S)) ** 2) <= 1).float()[None, None]
    coords = np.zeros((1, NT, 8), dtype=np.float64)
    coords[0, :, 0] = rs.randint(0, max(S - 64, 1), NT)
    coords[0, :, 1] = rs.randint(0, max(S - 64, 1), NT)
    coords[0, :, 2:4] = 64
    coords[0, :, 4] = 32
    coords[0, :, 5] = 1
    coords[0, :, 6] = rs.randint(0, 32, NT)
    coords[0, :, 7] = rs.randint(0, 32, NT)
    one = lambda v: torch.tensor([v])
    aug = dict(H=one(S), W=one(S), scale_factor_h=one(1), scale_factor_w=one(1), crop_size_h=one(S),
               crop_size_w=one(S), resize_ratio=one(1), crop_pos_x=one(0), crop_pos_y=one(0),
               resize_ratio_w=one(1.0), resize_ratio_h=one(1.0), patch_crop_size=one(32))
    return {
        "S": torch.rand(1, 1, S, S, generator=g) * 2 - 1,
        "I": torch.rand(1, 3, S, S, generator=g) * 2 - 1,
        "M": M,
        "T_images": torch.rand(1, NT, 2, 32, 32, generator=g),
        "val_T_images": torch.rand(1, NT, 2, 32, 32, generator=g),
        "I_masks": torch.ones(1, NT, 32, 32),
        "val_I_masks": torch.ones(1, NT, 32, 32),
        "T_coords": torch.from_numpy(coords),
        "val_T_coords": torch.from_numpy(coords.copy()),
        "augmentation_params": aug,
        "name": ["syn"], "S_paths": ["syn.png"], "full_T_coords": [],
    }


def step_inputs_from_batch(b, device=None):
    """What train_step() reads, from the set_input-style dict (`device`: move the tensors there, e.g. 'cuda' for the
    eager-PyTorch-on-B200 baseline arm of bench.py)."""
    NT = b["T_images"].shape[1]
    d = dict(real_S=b["S"], real_I=b["I"], M=b["M"],
             T_images=b["T_images"].reshape(NT, 2, 32, 32).float(),
             I_masks=b["I_masks"].reshape(NT, 1, 32, 32).float())
    if b.get("style_code") is not None:
        d["style_code"] = torch.as_tensor(b["style_code"]).float()
    if device is not None:
        d = {k: v.to(device) for k, v in d.items()}
    d["T_coords"] = b["T_coords"].numpy()
    return d


# --------------------------------------------------------------------------- evaluation metrics (models/model_utils.py:431-561)
def normal_angle_error_deg(real_T, fake_T, scale_nz=1.0):
    """compute_surface_normal_angle_error(compute_normal(real), compute_normal(fake), mode='evaluate').mean()
    (models/normal_losses.py:10-33 on models/model_utils.py:418-425): cosine similarity (eps 1e-6) of the unit normals, clamped,
    acos in degrees.  Pinned by tests/golden/metrics.npz (the reference's own functions)."""
    rn, fn = compute_normal(real_T, scale_nz), compute_normal(fake_T, scale_nz)
    c = F.cosine_similarity(fn, rn, dim=1, eps=1e-6).clamp(-1.0, 1.0)
    return (torch.acos(c) * 180.0 / math.pi).mean()


def _gauss11(sigma=1.5):
    d = torch.arange(11, dtype=torch.float32) - 5
    g = torch.exp(-(d / sigma) ** 2 / 2)
    return g / g.sum()


def ssim_torchmetrics(preds, target, data_range=1.0, k1=0.01, k2=0.03):
    """torchmetrics.functional.structural_similarity_index_measure(preds, target, data_range=...) with its defaults (Gaussian
    11x11, sigma 1.5, reduction 'elementwise_mean'), restated from torchmetrics 0.11 functional/image/ssim.py: reflect pad 5,
    depthwise valid conv of (p, t, p*p, t*t, p*t), SSIM index map, its 5-pixel border dropped, mean.  torchmetrics is a pip
    dependency of the reference (requirements.txt:18, unpinned) that is not installed here: PARITY UNPINNED."""
    c1, c2 = (k1 * data_range) ** 2, (k2 * data_range) ** 2
    B, C = preds.shape[:2]
    g = _gauss11().to(preds)
    k = (g[:, None] * g[None, :])[None, None].repeat(C, 1, 1, 1)
    p, t = F.pad(preds, (5, 5, 5, 5), mode="reflect"), F.pad(target, (5, 5, 5, 5), mode="reflect")
    o = F.conv2d(torch.cat([p, t, p * p, t * t, p * t]), k, groups=C)
    mp, mt, pp_, tt, pt = o.split(B)
    sp, st_, spt = pp_ - mp * mp, tt - mt * mt, pt - mp * mt
    idx = ((2 * mp * mt + c1) * (2 * spt + c2)) / ((mp * mp + mt * mt + c1) * (sp + st_ + c2))
    return idx[..., 5:-5, 5:-5].reshape(B, -1).mean(-1).mean()


def psnr_torchmetrics(preds, target, data_range=1.0):
    """torchmetrics.functional.peak_signal_noise_ratio(preds, target, data_range): 10 log10(data_range^2 / mean squared error).
    PARITY UNPINNED (see ssim_torchmetrics)."""
    return 10.0 * torch.log10(data_range ** 2 / ((preds - target) ** 2).mean())


def evaluation_metrics(real_I, fake_I, real_T, fake_T):
    """The I_PSNR / I_SSIM / T_AE / T_MSE part of compute_evaluation_metric (models/model_utils.py:483-498, 519-555)."""
    lo, hi = real_I.min(), real_I.max()
    r = (real_I - lo) / (hi - lo)
    f = ((fake_I - lo) / (hi - lo)).clamp(0, 1)
    fT = fake_T.clamp(0, 1)
    return dict(I_PSNR=psnr_torchmetrics(r, f).item(), I_SSIM=ssim_torchmetrics(r, f).item(),
                T_AE=normal_angle_error_deg(real_T, fT, 1.0).item(), T_MSE=((real_T - fT) ** 2).mean().item())


# --------------------------------------------------------------------------- stand-alone initial weights
def _xavier(shape, gen, gain=0.02):
    """init.xavier_normal_(w, gain) for a conv / linear weight (models/networks.py:204-222: gain 0.02, bias 0)."""
    rf = int(np.prod(shape[2:])) if len(shape) > 2 else 1
    std = gain * math.sqrt(2.0 / float(shape[1] * rf + shape[0] * rf))
    return torch.randn(*shape, generator=gen) * std


def init_resnet_g(input_nc=9, output_nc=5, ngf=64, n_blocks=9, seed=0):
    """A state_dict with the keys / shapes of define_G(..., 'resnet_9blocks', norm='instance') (models/networks.py:1057-1129;
    index map SURVEY.md A.2), initialised like init_weights('xavier', 0.02).  Lets bench.py's reference arm build its weights
    without importing the product package; tests/test_oracle_golden.py checks keys and shapes against the reference golden."""
    g = torch.Generator().manual_seed(seed)
    sd = {}

    def conv(name, co, ci, k):
        sd[name + ".weight"] = _xavier((co, ci, k, k), g)
        sd[name + ".bias"] = torch.zeros(co)

    def filt(name, c, taps, up):
        a = torch.tensor(taps)
        f = a[:, None] * a[None, :]
        f = f / f.sum() * (4.0 if up else 1.0)
        sd[name + ".filt"] = f[None, None].repeat(c, 1, 1, 1)

    conv("model.1", ngf, input_nc, 7)
    conv("model.4", ngf * 2, ngf, 3)
    filt("model.7", ngf * 2, [1.0, 2.0, 1.0], False)
    conv("model.8", ngf * 4, ngf * 2, 3)
    filt("model.11", ngf * 4, [1.0, 2.0, 1.0], False)
    for b in range(n_blocks):
        conv("model.%d.conv_block.1" % (12 + b), ngf * 4, ngf * 4, 3)
        conv("model.%d.conv_block.5" % (12 + b), ngf * 4, ngf * 4, 3)
    m = 12 + n_blocks
    filt("model.%d" % m, ngf * 4, [1.0, 3.0, 3.0, 1.0], True)
    conv("model.%d" % (m + 1), ngf * 2, ngf * 4, 3)
    filt("model.%d" % (m + 4), ngf * 2, [1.0, 3.0, 3.0, 1.0], True)
    conv("model.%d" % (m + 5), ngf, ngf * 2, 3)
    conv("model.%d" % (m + 9), output_nc, ngf, 7)
    return sd


def init_unet_custom(input_nc=9, ngf=10, num_downs=8, num_layer_separate=4, seed=0):
    """A state_dict with the keys / shapes of define_G(..., 'unet256_custom', norm='instance') — the reference's default generator
    (models/networks.py:1430-1573; key map SURVEY.md A.1: 40 tensors at ngf 10), initialised like init_weights('xavier', 0.02).
    ConvTranspose2d weights are [Cin, Cout, 4, 4]; levels num_layer_separate-1 .. 0 have twin RGB / touch decoders."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    ch = [ngf * min(2 ** i, 8) for i in range(num_downs)]          # 10, 20, 40, 80, 80, 80, 80, 80
    cin = input_nc
    for i in range(num_downs):
        key = "down%d.model.%d" % (i, 0 if i == 0 else 1)
        sd[key + ".weight"] = _xavier((ch[i], cin, 4, 4), g)
        sd[key + ".bias"] = torch.zeros(ch[i])
        cin = ch[i]
    for i in range(num_downs - 1, -1, -1):
        c_in = ch[i] if i == num_downs - 1 else 2 * ch[i]           # innermost level has no skip concatenation
        outs = {"up%d" % i: (ch[i - 1] if i > 0 else 3)}
        if i < num_layer_separate:
            outs["up%d_T" % i] = ch[i - 1] if i > 0 else 2
        if i == 0:
            c_in = ch[0]        # up0 takes the level-1 decoder output only (networks.py:1567-1573; A.1: up0.model.1.weight (10, 3, 4, 4))
        for name, c_out in outs.items():
            sd[name + ".model.1.weight"] = _xavier((c_in, c_out, 4, 4), g)
            sd[name + ".model.1.bias"] = torch.zeros(c_out)
    return sd


def init_multiscale_d(input_nc, ndf=64, n_layers=3, num_D=3, seed=0):
    """State_dict of define_D(..., 'multiscale', norm='batch') (models/networks.py:1649-1750; index map SURVEY.md A.5):
    conv weights xavier(0.02), biases 0, BatchNorm weight ~ N(1, 0.02), bias 0 (:223-226), fresh running statistics."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for d in range(num_D):
        pre = "layer%d." % d

        def conv(i, co, ci):
            sd["%s%d.weight" % (pre, i)] = _xavier((co, ci, 4, 4), g)
            sd["%s%d.bias" % (pre, i)] = torch.zeros(co)

        def bn(i, c):
            sd["%s%d.weight" % (pre, i)] = 1.0 + 0.02 * torch.randn(c, generator=g)
            sd["%s%d.bias" % (pre, i)] = torch.zeros(c)
            sd["%s%d.running_mean" % (pre, i)] = torch.zeros(c)
            sd["%s%d.running_var" % (pre, i)] = torch.ones(c)
            sd["%s%d.num_batches_tracked" % (pre, i)] = torch.tensor(0, dtype=torch.long)

        conv(0, ndf, input_nc)
        nf, i = ndf, 2
        for _ in range(1, n_layers):
            nf_prev, nf = nf, min(nf * 2, 512)
            conv(i, nf, nf_prev)
            bn(i + 1, nf)
            i += 3
        nf_prev, nf = nf, min(nf * 2, 512)
        conv(i, nf, nf_prev)
        bn(i + 1, nf)
        conv(i + 3, 1, nf)
    return sd
