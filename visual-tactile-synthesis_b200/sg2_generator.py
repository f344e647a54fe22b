"""`--netG stylegan2 | smallstylegan2` on the B200 path: forward (inference / feature taps) of the reference's
StyleGAN2Generator (models/stylegan_networks.py:912-929; encoder :800-851, decoder :854-909).

Parameter tree and state_dict keys are the reference's (`encoder.convs.2.conv2.1.weight`, `decoder.convs.3.conv.weight`
[1,co,ci,3,3], the `kernel` blur buffers, ...), so its checkpoints load.  The compute is not a translation:

* every blur is folded into the neighbouring strided / transposed conv on the device (`skit_sg2_weight_prep`), so the
  whole net is plain zero-padded convolutions: 3x3 s1, 6x6 s2 (blur + 3x3 s2), 4x4 s2 (blur + 1x1 s2 skip), and one
  3x3 conv to 4*co sub-pixel channels (conv_transpose + blur) followed by depth-to-space — all of them shapes the
  tcgen05 conv kernels take when the widths are multiples of 64 (ngf 64: 64/128 channels);
* FusedLeakyReLU, NoiseInjection, the ResBlock merge, depth-to-space, zero halo and the bf16 hi/lo split of the next
  conv's operand are one element-wise pass (`skit_sg2_bias_act`).

`fwd` / `bwd` are the explicit training pair (every parameter's gradient, through the transposed fold); the skitG train
step itself is not wired to this generator (3 output channels, stylegan_networks.py:892): `SinSKITGModel(netG='stylegan2')` raises.
"""
import math

import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .ops import FMT_BF16X2, FMT_F32, PAD_ZERO, _p
from .networks import _FlatParamsMixin

SQRT2 = math.sqrt(2.0)


def channel_table(ngf):
    """stylegan_networks.py:805-816 (encoder) == :862-873 (decoder)."""
    m = ngf / 32
    t = {4: min(512, int(round(4096 * m))), 8: min(512, int(round(2048 * m))), 16: min(512, int(round(1024 * m))),
         32: min(512, int(round(512 * m)))}
    t.update({64: int(round(256 * m)), 128: int(round(128 * m)), 256: int(round(64 * m)), 512: int(round(32 * m)),
              1024: int(round(16 * m))})
    return t


def _blur_kernel(gain=1.0):
    k = torch.tensor([1.0, 3.0, 3.0, 1.0])
    k = k[None, :] * k[:, None]
    return k / k.sum() * gain


# ----------------------------------------------------------------------------- parameter holders (reference key names)
class EqualConv2d(nn.Module):
    """weight [co,ci,k,k] ~ N(0,1), used as weight / sqrt(ci k k) (stylegan_networks.py:159-190); bias only without activation."""

    def __init__(self, ci, co, k, stride=1, bias=False):
        super().__init__()
        self.ci, self.co, self.k, self.stride = ci, co, k, stride
        self.weight = nn.Parameter(torch.randn(co, ci, k, k))
        self.bias = nn.Parameter(torch.zeros(co)) if bias else None


class FusedLeakyReLU(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.bias = nn.Parameter(torch.zeros(1, c, 1, 1))


class Blur(nn.Module):
    """Holds the reference's `kernel` buffer only (state_dict compatibility); the blur itself lives in the folded filters."""

    def __init__(self, gain=1.0):
        super().__init__()
        self.register_buffer("kernel", _blur_kernel(gain))


class ConvLayer(nn.Sequential):
    """[Blur] + EqualConv2d [+ FusedLeakyReLU] with the reference's Sequential indices (stylegan_networks.py:613-665)."""

    def __init__(self, ci, co, k, downsample=False, activate=True):
        mods = []
        if downsample:
            mods.append(Blur())
        mods.append(EqualConv2d(ci, co, k, stride=2 if downsample else 1, bias=False))
        if activate:
            mods.append(FusedLeakyReLU(co))
        super().__init__(*mods)
        self.downsample, self.activate, self.k, self.ci, self.co = downsample, activate, k, ci, co

    @property
    def conv(self):
        return self[1] if self.downsample else self[0]

    @property
    def act_bias(self):
        return self[-1].bias if self.activate else None


class ResBlock(nn.Module):
    def __init__(self, ci, co, downsample=True):
        super().__init__()
        self.ci, self.co, self.downsample = ci, co, downsample
        self.conv1 = ConvLayer(ci, ci, 3)
        self.conv2 = ConvLayer(ci, co, 3, downsample=downsample)
        if ci != co or downsample:
            self.skip = ConvLayer(ci, co, 1, downsample=downsample, activate=False)
        else:
            self.skip = nn.Identity()


class ModulatedConv2d(nn.Module):
    """style_dim=None, demodulate=True, upsample=True (the only configuration the generator builds, :886-889)."""

    def __init__(self, ci, co, k):
        super().__init__()
        self.ci, self.co, self.k = ci, co, k
        self.weight = nn.Parameter(torch.randn(1, co, ci, k, k))
        self.blur = Blur(gain=4.0)


class NoiseInjection(nn.Module):
    def __init__(self):
        super().__init__()
        self.weight = nn.Parameter(torch.zeros(1))


class StyledConv(nn.Module):
    def __init__(self, ci, co, k, inject_noise=True):
        super().__init__()
        self.inject_noise = inject_noise
        self.conv = ModulatedConv2d(ci, co, k)
        self.noise = NoiseInjection()
        self.activate = FusedLeakyReLU(co)


class _Convs(nn.Module):
    def __init__(self, mods):
        super().__init__()
        self.convs = nn.Sequential(*mods)


# ----------------------------------------------------------------------------- effective (folded) filters
class _EffConv:
    """One launchable conv: the folded filter tensor built by skit_sg2_weight_prep, its packs, and (for training) the
    gradient w.r.t. the folded filter, folded back into the parameter's gradient by skit_sg2_weight_prep_bwd."""

    def __init__(self, src, mode, stride, pad, co_pad=0):
        w = src.weight
        co, ci, k = (w.shape[1], w.shape[2], w.shape[3]) if w.dim() == 5 else (w.shape[0], w.shape[1], w.shape[2])
        self.src, self.mode, self.stride, self.pad = src, mode, stride, pad
        self.ci, self.co_src, self.k_src = ci, co, k
        self.k = k + 3 if mode == 1 else k
        self.co = 4 * co if mode == 2 else max(co, co_pad)
        self.weight = torch.zeros(self.co, ci, self.k, self.k, dtype=torch.float32, device=w.device)
        self.use_tc = ops_tc_enabled() and ci % 64 == 0 and self.co % 64 == 0 and (stride == 1 or self.k % 2 == 0)
        self.packs = {}
        self.grad = None

    def pack_mode(self, mode):
        """0 forward, 1 stride-1 tcgen05 input gradient, 2 gather input gradient (CUDA cores), 3 stride-2 tcgen05 input gradient."""
        pk = self.packs.get(mode)
        if pk is None:
            pk = self.packs[mode] = ops.PackedWeights(self.weight, mode, want_f32=not self.use_tc, want_bf16=self.use_tc)
        return pk

    def refresh(self):
        L.call("skit_sg2_weight_prep", _p(self.src.weight), self.co_src, self.ci, self.k_src, self.mode, _p(self.weight), L.stream())
        if not self.packs:
            self.pack_mode(0)
        else:
            for pk in self.packs.values():
                pk.refresh(self.weight)

    def __call__(self, x_op, ho, wo):
        y, _ = ops.conv2d_fwd(x_op, self.pack_mode(0), self.stride, x_op.pad - self.pad, ho, wo)
        return y

    # -- backward
    @property
    def q(self):
        """Zero halo the raw-output gradient operand needs for this conv's input-gradient kernel."""
        if not self.use_tc:
            return 0
        return self.k - 1 if self.stride == 1 else self.k // 2 - 1

    @property
    def dfmt(self):
        return FMT_BF16X2 if self.use_tc else FMT_F32

    def bwd(self, x_op, d_op, ho, wo, need_dgrad=True):
        """d_op: gradient w.r.t. the raw output (halo self.q).  Accumulates the folded filter's gradient; returns the gradient
        w.r.t. the haloed input operand [n][hp][wp][ci] (fp32)."""
        if self.grad is None:
            self.grad = torch.zeros_like(self.weight)
        ops.conv2d_wgrad(x_op, x_op.pad - self.pad, d_op, self.q, self.k, self.stride, ho, wo, self.grad)
        if not need_dgrad:
            return None
        assert x_op.pad == self.pad
        if self.use_tc and self.stride == 1:
            return ops.conv2d_dgrad_s1(d_op, self.pack_mode(1))
        if self.use_tc:
            return ops.conv2d_dgrad_s2(d_op, self.q, self.pack_mode(3), self.k, ho, wo, x_op.hp, x_op.wp)
        return ops.conv2d_dgrad_gather(d_op.data, self.pack_mode(2), self.stride, x_op.hp, x_op.wp)

    def fold_grad(self):
        """param.grad += (d folded / d param)^T grad, then clear."""
        if self.grad is None:
            return
        w = self.src.weight
        if w.grad is None:
            w.grad = torch.zeros_like(w)
        L.call("skit_sg2_weight_prep_bwd", _p(w), self.co_src, self.ci, self.k_src, self.mode, _p(self.grad), _p(w.grad), L.stream())
        self.grad.zero_()


def _grad(p):
    """p.grad, allocated (zeros) on first use."""
    if p.grad is None:
        p.grad = torch.zeros_like(p)
    return p.grad


def ops_tc_enabled():
    from . import networks
    return networks.TC_ENABLED


def bias_act(raw, c, bias=None, noise=None, noise_w=None, skip=None, shuffle=False, act=True, gain=SQRT2, post=1.0,
             want_dense=False, op_pad=None, op_fmt=FMT_F32, nchw_c=0):
    """skit_sg2_bias_act -> (dense NHWC or None, Operand or None, NCHW or None)."""
    n, h, w, craw = raw.shape
    H, W = (2 * h, 2 * w) if shuffle else (h, w)
    dense = torch.empty((n, H, W, c), dtype=torch.float32, device=raw.device) if want_dense else None
    op = ops.Operand(n, H, W, c, op_pad, op_fmt, raw.device) if op_pad is not None else None
    nchw = torch.empty((n, nchw_c, H, W), dtype=torch.float32, device=raw.device) if nchw_c else None
    L.call("skit_sg2_bias_act", _p(raw), n, h, w, craw, c, _p(bias), _p(noise), _p(noise_w) if noise is not None else None,
           _p(skip), int(shuffle), int(act), gain, post, _p(dense), op.ref() if op is not None else None,
           op_pad or 0, _p(nchw), nchw_c, L.stream())
    return dense, op, nchw


def bias_act_bwd(raw, n, h, w, c, eff, bias=None, noise=None, noise_w=None, shuffle=False, act=True, gain=SQRT2, post=1.0,
                 dpa=None, pa=0, dpb=None, pb=0, dd=None, want_dskip=False, dbias=None, dnoise_w=None):
    """skit_sg2_bias_act_bwd -> (gradient operand for `eff`'s backward, dskip or None).  h, w: raw resolution."""
    dev = (dpa if dpa is not None else dd if dd is not None else dpb).device
    ce = 4 * c if shuffle else c
    d_op = ops.Operand(n, h, w, ce, eff.q, eff.dfmt, dev)
    H, W = (2 * h, 2 * w) if shuffle else (h, w)
    dskip = torch.empty((n, H, W, c), dtype=torch.float32, device=dev) if want_dskip else None
    L.call("skit_sg2_bias_act_bwd", _p(raw), n, h, w, raw.shape[3] if raw is not None else ce, c, _p(bias), _p(noise),
           _p(noise_w) if noise is not None else None, int(shuffle), int(act), gain, post, _p(dpa), pa, _p(dpb), pb, _p(dd),
           d_op.ref(), eff.q, _p(dskip), _p(dbias), _p(dnoise_w) if noise is not None else None, L.stream())
    return d_op, dskip


# ----------------------------------------------------------------------------- the generator
class StyleGAN2Generator(_FlatParamsMixin, nn.Module):
    """define_G('stylegan2' | 'smallstylegan2') (networks.py:307-310).  forward(input, layers=[], encode_only=False)."""

    def __init__(self, input_nc, output_nc, ngf=64, use_dropout=False, n_blocks=6, opt=None, **unused):
        super().__init__()
        assert opt is not None
        self.opt, self.input_nc, self.n_blocks = opt, input_nc, n_blocks
        ch = channel_table(ngf)
        res = 2 ** int(round(math.log2(min(opt.load_size, opt.crop_size))))   # np.rint(np.log2(...)) in the reference (:820)
        nd = self.num_down = opt.stylegan2_G_num_downsampling
        enc = [nn.Identity(), ConvLayer(input_nc, ch[res], 1)]     # a KeyError for 2048 (crop 1536) like the reference
        cur = res
        for _ in range(nd):
            enc.append(ResBlock(ch[cur], ch[cur // 2], downsample=True))
            cur //= 2
        for _ in range(n_blocks // 2):
            enc.append(ResBlock(ch[cur], ch[cur], downsample=False))
        dec = [ResBlock(ch[cur], ch[cur], downsample=False) for _ in range(n_blocks // 2)]
        inject = "small" not in opt.netG
        for _ in range(nd):
            dec.append(StyledConv(ch[cur], ch[cur * 2], 3, inject_noise=inject))
            cur *= 2
        dec.append(ConvLayer(ch[cur], 3, 1))
        self.encoder, self.decoder = _Convs(enc), _Convs(dec)
        self.__dict__["_eff"] = None
        self.__dict__["_eff_key"] = None

    # -- folded filters: rebuilt when any parameter changed (in-place version counters) or moved
    def refresh_packs(self):
        params = list(self.parameters())
        key = tuple((p.data_ptr(), p._version) for p in params)
        if self._eff is not None and key == self._eff_key:
            return
        if any(c % 4 for c in self._widths()):
            raise NotImplementedError("B200 path: StyleGAN2Generator needs channel widths that are multiples of 4 "
                                      "(ngf a multiple of 2), got %s" % sorted(set(self._widths())))
        eff = {}

        def layer(cl):        # ConvLayer -> _EffConv
            if cl.downsample:
                return _EffConv(cl.conv, 1, 2, 2 if cl.k == 3 else 1)
            return _EffConv(cl.conv, 0, 1, cl.k // 2, co_pad=4 if cl.co < 4 else 0)

        for m in list(self.encoder.convs) + list(self.decoder.convs):
            if isinstance(m, ConvLayer):
                eff[m] = layer(m)
            elif isinstance(m, ResBlock):
                eff[m.conv1], eff[m.conv2] = layer(m.conv1), layer(m.conv2)
                if isinstance(m.skip, ConvLayer):
                    eff[m.skip] = layer(m.skip)
            elif isinstance(m, StyledConv):
                eff[m] = _EffConv(m.conv, 2, 1, 1)
        for e in eff.values():
            e.refresh()
        last = self.decoder.convs[-1]
        b = torch.zeros(4, dtype=torch.float32, device=params[0].device)   # the 3-channel head runs zero-padded to 4 channels
        b[:3] = last.act_bias.detach().reshape(-1)
        self.__dict__["_head_bias"] = b
        self.__dict__["_eff"], self.__dict__["_eff_key"] = eff, key

    def _widths(self):
        out = []
        for m in self.modules():
            if isinstance(m, EqualConv2d) and m.co != 3:
                out += [m.co]
            if isinstance(m, ModulatedConv2d):
                out += [m.ci, m.co]
        return out

    # -- pieces
    def _fmt(self, *consumers):
        return FMT_BF16X2 if any(self._eff[c].use_tc for c in consumers) else FMT_F32

    def _resblock(self, m, x_op, x_dense, h, w, nxt, saved=None):
        """x_op: zero-haloed (pad 1) operand of the block input; returns (dense, operand pad 1 in the format `nxt` wants)."""
        E = self._eff
        raw1 = E[m.conv1](x_op, h, w)
        if m.downsample:
            _, h1, _ = bias_act(raw1, m.ci, bias=m.conv1.act_bias, op_pad=2, op_fmt=self._fmt(m.conv2))
            ho, wo = h // 2, w // 2
            raw2 = E[m.conv2](h1, ho, wo)
            skip = E[m.skip](x_op, ho, wo)                      # blur + 1x1 s2 == 4x4 s2 pad 1 over the same operand
        else:
            _, h1, _ = bias_act(raw1, m.ci, bias=m.conv1.act_bias, op_pad=1, op_fmt=self._fmt(m.conv2))
            ho, wo = h, w
            raw2 = E[m.conv2](h1, ho, wo)
            skip = E[m.skip](x_op, ho, wo) if isinstance(m.skip, ConvLayer) else x_dense
        dense, op, _ = bias_act(raw2, m.co, bias=m.conv2.act_bias, skip=skip, gain=SQRT2, post=1.0 / SQRT2,
                                want_dense=True, op_pad=1, op_fmt=nxt)
        if saved is not None:
            saved.append(("res", m, x_op, h, w, raw1, h1, raw2, ho, wo))
        return dense, op, ho, wo

    def _resblock_bwd(self, rec, grads):
        """grads = (dpa, pa, dpb, pb, dd): gradient w.r.t. the block output as haloed tensors and / or dense.  Returns the
        same tuple for the block input."""
        _, m, x_op, h, w, raw1, h1, raw2, ho, wo = rec
        E = self._eff
        e1, e2 = E[m.conv1], E[m.conv2]
        n = x_op.n
        dpa, pa, dpb, pb, dd = grads
        _grad(m.conv2.act_bias), _grad(m.conv1.act_bias)
        d2, dskip = bias_act_bwd(raw2, n, ho, wo, m.co, e2, bias=m.conv2.act_bias, gain=SQRT2, post=1.0 / SQRT2, dpa=dpa, pa=pa,
                                 dpb=dpb, pb=pb, dd=dd, want_dskip=True, dbias=m.conv2.act_bias.grad)
        dp_h1 = e2.bwd(h1, d2, ho, wo)
        d1, _ = bias_act_bwd(raw1, n, h, w, m.ci, e1, bias=m.conv1.act_bias, dpa=dp_h1, pa=h1.pad, dbias=m.conv1.act_bias.grad)
        dp_x = e1.bwd(x_op, d1, h, w)
        if isinstance(m.skip, ConvLayer):
            es = E[m.skip]
            ds, _ = bias_act_bwd(None, n, ho, wo, m.co, es, act=False, dd=dskip)     # dense -> the skip conv's gradient operand
            return dp_x, x_op.pad, es.bwd(x_op, ds, ho, wo), x_op.pad, None
        return dp_x, x_op.pad, None, 0, dskip

    def _next_fmt(self, seq, i):
        """Operand format the consumer(s) of block i's output want."""
        if i + 1 >= len(seq):
            return FMT_F32
        m = seq[i + 1]
        if isinstance(m, ResBlock):
            cons = [m.conv1] + ([m.skip] if isinstance(m.skip, ConvLayer) else [])
            return self._fmt(*cons)
        return self._fmt(m)

    def forward(self, input, layers=[], encode_only=False, noises=None):
        """input NCHW fp32 -> fake [n,3,S,S] (+ encoder features at `layers`, indices into encoder.convs).
        noises: optional list of [n,1,H,W] tensors, one per StyledConv (default: drawn on the device like NoiseInjection)."""
        fake, feats, _ = self.fwd(input, layers=layers, encode_only=encode_only, noises=noises, save=False)
        if encode_only:
            return feats
        return (fake, feats) if len(layers) > 0 else fake

    def fwd(self, input, layers=(), encode_only=False, noises=None, save=True):
        """Explicit forward -> (fake NCHW or None, encoder features, ctx for `bwd` when save)."""
        if not input.is_cuda:
            raise RuntimeError("StyleGAN2Generator runs only on a CUDA device through libskit_b200.so; there is no CPU fallback")
        self.refresh_packs()
        E = self._eff
        layers = list(layers)
        enc, dec = list(self.encoder.convs), list(self.decoder.convs)
        if -1 in layers:
            layers.append(len(enc) - 1)
        x = input.contiguous().float()
        n, _, h, w = x.shape
        feats = []
        saved = [] if save else None
        if 0 in layers:
            feats.append(x)
        # stem: 1x1 conv on the raw input (thin: CUDA-core fp32 conv), activation writes the first block's operand
        x_op = ops.nchw_cat_to_operand([x], 0, PAD_ZERO)
        raw = E[enc[1]](x_op, h, w)
        seq = enc + dec
        dense, op, _ = bias_act(raw, enc[1].co, bias=enc[1].act_bias, want_dense=True, op_pad=1, op_fmt=self._next_fmt(seq, 1))
        if save:
            saved.append(("stem", enc[1], x_op, h, w, raw))
        if 1 in layers:
            feats.append(dense.permute(0, 3, 1, 2))
        for i in range(2, len(enc)):
            # the encoder's last block feeds the decoder's first: one sequence
            dense, op, h, w = self._resblock(enc[i], op, dense, h, w, self._next_fmt(seq, i), saved)
            if i in layers:
                feats.append(dense.permute(0, 3, 1, 2))
        if encode_only:
            return None, feats, saved
        si = 0
        for j, m in enumerate(dec[:-1]):
            i = len(enc) + j
            if isinstance(m, ResBlock):
                dense, op, h, w = self._resblock(m, op, dense, h, w, self._next_fmt(seq, i), saved)
                continue
            raw = E[m](op, h, w)                                  # [n, h, w, 4*co]: 2x2 sub-pixels of the up-sampled map
            nz = None
            if m.inject_noise:
                nz = noises[si] if noises is not None else torch.randn(n, 1, 2 * h, 2 * w, device=x.device)
                nz = nz.contiguous().float()
                si += 1
            last = j + 1 == len(dec) - 1
            nxt = FMT_F32 if last else self._next_fmt(seq, i)
            if save:
                saved.append(("up", m, op, h, w, raw, nz))
            dense, op, _ = bias_act(raw, m.conv.co, bias=m.activate.bias, noise=nz, noise_w=m.noise.weight if nz is not None else None,
                                    shuffle=True, want_dense=not last, op_pad=0 if last else 1, op_fmt=nxt)
            h, w = 2 * h, 2 * w
        raw = E[dec[-1]](op, h, w)                                # 1x1 conv to 3 (+1 zero) channels
        if save:
            saved.append(("head", dec[-1], op, h, w, raw))
        _, _, fake = bias_act(raw, 4, bias=self._head_bias, nchw_c=3)
        return fake, feats, saved

    def bwd(self, saved, dfake):
        """dfake: gradient w.r.t. `fake` (NCHW [n,3,S,S]).  Accumulates into every parameter's .grad (weights through the
        transposed blur / scale / demodulation fold, FusedLeakyReLU biases, noise strengths)."""
        E = self._eff
        n = dfake.shape[0]
        grads = None
        for rec in reversed(saved):
            kind = rec[0]
            if kind == "head":
                _, m, op, h, w, raw = rec
                dd = torch.zeros((n, h, w, 4), dtype=torch.float32, device=dfake.device)
                dd[..., :3] = dfake.permute(0, 2, 3, 1)
                db = torch.zeros(4, dtype=torch.float32, device=dfake.device)
                d, _ = bias_act_bwd(raw, n, h, w, 4, E[m], bias=self._head_bias, dd=dd, dbias=db)
                dp = E[m].bwd(op, d, h, w)
                _grad(m.act_bias).view(-1).add_(db[:3])
                grads = (dp, op.pad, None, 0, None)
            elif kind == "up":
                _, m, op, h, w, raw, nz = rec
                dpa, pa, dpb, pb, dd = grads
                d, _ = bias_act_bwd(raw, n, h, w, m.conv.co, E[m], bias=m.activate.bias, noise=nz,
                                    noise_w=m.noise.weight if nz is not None else None, shuffle=True, dpa=dpa, pa=pa, dpb=dpb, pb=pb,
                                    dd=dd, dbias=_grad(m.activate.bias), dnoise_w=_grad(m.noise.weight) if nz is not None else None)
                grads = (E[m].bwd(op, d, h, w), op.pad, None, 0, None)
            elif kind == "res":
                grads = self._resblock_bwd(rec, grads)
            else:
                _, m, x_op, h, w, raw = rec
                dpa, pa, dpb, pb, dd = grads
                d, _ = bias_act_bwd(raw, n, h, w, m.co, E[m], bias=m.act_bias, dpa=dpa, pa=pa, dpb=dpb, pb=pb, dd=dd, dbias=_grad(m.act_bias))
                E[m].bwd(x_op, d, h, w, need_dgrad=False)
        for e in E.values():
            e.fold_grad()
