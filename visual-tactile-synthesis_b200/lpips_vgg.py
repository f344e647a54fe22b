"""LPIPS-VGG16 perceptual loss on the B200 path.

The reference builds `lpips.LPIPS(net="vgg")` (pip lpips 0.1.4; models/sinskitG_model.py:495) and back-propagates it into
the generator from the full-resolution image (`:1711`, weight `lambda_G1_lpips`, default 1) and from the 32x32 touch
patches, one channel at a time (`:1619-1658`, weight `lambda_G2_lpips`, default 10) — SURVEY §8(a11): it dominates the
default step.  Here the frozen VGG16 trunk runs on the library's conv kernels (tcgen05 for every layer but the 3-channel
stem), forward for both images and input-gradient only for the generated one; ScalingLayer, MaxPool2d and the LPIPS head
are `csrc/lpips_ops.cu`.

state_dict keys are the package's (`net.slice3.12.weight`, `lin2.model.1.weight`, `lins.2.model.1.weight`,
`scaling_layer.shift`), so `lpips`' own checkpoints load; without them (no network here) the weights are random.
"""
import torch
import torch.nn as nn

from . import _lib as L
from . import ops
from .networks import (ACT_RELU, Conv2d, FMT_BF16X2, FMT_F32, NORM_NONE, PAD_ZERO, _conv_fwd, _FlatParamsMixin, _Placeholder,
                       _stage_bwd, _wgrad_join)
from .ops import _p

VGG_SLICES = ((0, 2), (5, 7), (10, 12, 14), (17, 19, 21), (24, 26, 28))   # conv indices inside torchvision's vgg16.features
VGG_CHNS = (64, 128, 256, 512, 512)


class _Lin1x1(nn.Module):
    def __init__(self, c):
        super().__init__()
        self.weight = nn.Parameter(torch.rand(1, c, 1, 1) / c * 4)


class _NetLinLayer(nn.Module):
    """lpips.NetLinLayer: Sequential(Dropout, Conv2d(c, 1, 1, bias=False)); eval mode, so the dropout is the identity."""

    def __init__(self, c):
        super().__init__()
        self.model = nn.Sequential(_Placeholder("Dropout"), _Lin1x1(c))


class _ScalingLayer(nn.Module):
    def __init__(self):
        super().__init__()
        self.register_buffer("shift", torch.tensor([-0.030, -0.088, -0.188])[None, :, None, None])
        self.register_buffer("scale", torch.tensor([0.458, 0.448, 0.450])[None, :, None, None])


class _Vgg16(nn.Module):
    def __init__(self):
        super().__init__()
        cin = 3
        for s, idxs in enumerate(VGG_SLICES):
            sl = nn.Module()
            for i in idxs:
                sl.add_module(str(i), Conv2d(cin, VGG_CHNS[s], 3, bias=True))
                cin = VGG_CHNS[s]
            self.add_module("slice%d" % (s + 1), sl)

    def slices(self):
        return [[getattr(getattr(self, "slice%d" % (s + 1)), str(i)) for i in idxs] for s, idxs in enumerate(VGG_SLICES)]


class LPIPS(_FlatParamsMixin, nn.Module):
    """`lpips.LPIPS(net='vgg')` with an explicit value-and-input-gradient entry point (`loss_and_grad`)."""

    def __init__(self, net="vgg", **unused):
        super().__init__()
        if net != "vgg":
            raise NotImplementedError("B200 path: LPIPS is built for net='vgg' (the training criterion, sinskitG_model.py:495); got %r" % net)
        self.scaling_layer = _ScalingLayer()
        self.net = _Vgg16()
        for k, c in enumerate(VGG_CHNS):
            setattr(self, "lin%d" % k, _NetLinLayer(c))
        self.lins = nn.ModuleList([getattr(self, "lin%d" % k) for k in range(5)])
        with torch.no_grad():     # no pretrained checkpoint offline: He-normal trunk so that random-weight features stay O(1)
            for convs in self.net.slices():
                for c in convs:
                    nn.init.kaiming_normal_(c.weight, nonlinearity="relu")
                    c.bias.normal_(0, 0.05)
        for p in self.parameters():
            p.requires_grad_(False)
        self.eval()

    # -- VGG16 trunk.  x: NCHW fp32 with 1 or 3 channels, sides multiples of 16.
    def fwd(self, x, save=False):
        n, cin, h, w = x.shape
        if cin not in (1, 3) or h % 16 or w % 16:
            raise ValueError("LPIPS: input must have 1 or 3 channels and sides that are multiples of 16, got %s" % (tuple(x.shape),))
        dev = x.device
        op = ops.Operand(n, h, w, 3, 1, FMT_F32, dev)
        L.call("skit_lpips_scale_fwd", _p(x), n, cin, h, w, op.ref(), L.stream())
        feats, saved = [], []
        f = None

        def fmt(conv):
            return FMT_BF16X2 if conv.use_tc else FMT_F32

        for s, convs in enumerate(self.net.slices()):
            if s > 0:
                c = f.shape[3]
                h, w = h // 2, w // 2
                op = ops.Operand(n, h, w, c, 1, fmt(convs[0]), dev)
                L.call("skit_maxpool2_fwd", _p(f), n, 2 * h, 2 * w, c, op.ref(), 1, L.stream())
            for j, conv in enumerate(convs):
                raw, _ = _conv_fwd(conv, op, 0, h, w, NORM_NONE)
                if save:
                    saved.append((conv, op, raw))
                if j + 1 == len(convs):
                    f, _ = ops.norm_act_pad(raw, act=ACT_RELU, want_dense=True)
                    feats.append(f)
                else:
                    _, op = ops.norm_act_pad(raw, act=ACT_RELU, pad=1, pad_mode=PAD_ZERO, fmt=fmt(convs[j + 1]))
        return (feats, saved) if save else feats

    def bwd(self, feats, saved, dfeats):
        """dfeats: gradients w.r.t. the five tapped (post-ReLU) feature maps -> gradient w.r.t. the stem's haloed operand."""
        dpad, dadd = None, dfeats[4]
        idx = len(saved)
        sl = self.net.slices()
        for s in range(4, -1, -1):
            for _ in range(len(sl[s])):
                idx -= 1
                conv, x_op, raw = saved[idx]
                dpad = _stage_bwd(conv, x_op, raw, None, NORM_NONE, ACT_RELU, raw.shape[1] * raw.shape[2], dpad=dpad, pad=1,
                                  pad_mode=PAD_ZERO, dadd=dadd, need_dgrad=True, need_wgrad=False)
                dadd = None
            if s > 0:
                fp = feats[s - 1]
                n, h, w, c = fp.shape
                df = torch.empty_like(fp)
                L.call("skit_maxpool2_bwd", _p(fp), _p(dpad), n, h, w, c, 1, _p(dfeats[s - 1]), _p(df), L.stream())
                dpad, dadd = None, df
        _wgrad_join()
        return dpad

    def loss_and_grad(self, fake, real, gscale=1.0, want_grad=True, dx=None, dx_c0=0, accumulate=False, real_feats=None):
        """LPIPS(fake, real) per sample [n] and gscale * d(sum_b LPIPS_b) / d fake, written (or accumulated) into channels
        [dx_c0, dx_c0 + fake.shape[1]) of the NCHW tensor `dx` (allocated when None)."""
        if not fake.is_cuda:
            raise RuntimeError("LPIPS runs only on a CUDA device through libskit_b200.so; there is no CPU fallback")
        self.refresh_packs_once()
        fake, real = fake.contiguous().float(), (real.contiguous().float() if real is not None else None)
        n, cin, h, w = fake.shape
        fr = real_feats if real_feats is not None else self.fwd(real)
        if want_grad:
            ff, saved = self.fwd(fake, save=True)
        else:
            ff, saved = self.fwd(fake), None
        loss = torch.zeros(n, dtype=torch.float32, device=fake.device)
        dfe = []
        for k in range(5):
            _, hk, wk, ck = ff[k].shape
            df = torch.empty_like(ff[k]) if want_grad else None
            lw = getattr(self, "lin%d" % k).model[1].weight
            L.call("skit_lpips_layer", _p(ff[k]), _p(fr[k]), _p(lw), n, hk, wk, ck, float(gscale), _p(loss), _p(df), L.stream())
            dfe.append(df)
        if not want_grad:
            return loss, None
        dop = self.bwd(ff, saved, dfe)
        if dx is None:
            dx = torch.empty((n, cin, h, w), dtype=torch.float32, device=fake.device)
            accumulate = False
        L.call("skit_lpips_scale_bwd", _p(dop), n, cin, h, w, 1.0, _p(dx), dx.shape[1], dx_c0, int(accumulate), L.stream())
        return loss, dx

    def refresh_packs_once(self):
        """The criterion is frozen: packs are built on first use and again only if a weight tensor was written in place."""
        key = tuple((p.data_ptr(), p._version) for p in self.parameters())
        if self.__dict__.get("_pack_key") != key:
            for convs in self.net.slices():
                for c in convs:
                    c.pack(0)
                    c.pack(1 if c.use_tc else 2)
            self.refresh_packs()
            self.__dict__["_pack_key"] = key

    # -- package API: LPIPS.forward(in0, in1) -> [N,1,1,1]  (normalize=False: inputs already in [-1, 1])
    def forward(self, in0, in1, retPerLayer=False, normalize=False):
        if normalize:
            in0, in1 = 2 * in0 - 1, 2 * in1 - 1
        loss, _ = self.loss_and_grad(in0, in1, want_grad=False)
        return loss.view(-1, 1, 1, 1)
