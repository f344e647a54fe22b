"""Drop-in mirror of the reference's `models/networks.py` API for the skitG/sinskitG hot path:
define_G / define_D / define_F / GANLoss / PatchSampleF / init_net / get_scheduler, with
identical signatures, error behaviour and state_dict key names — but
every forward and backward runs as explicit launches of the sm_100a kernels behind
include/skit_b200.h (no autograd, no ATen convolution on the hot path).

Reference: models/networks.py:148-174 (schedulers), :191-252 (init), :255-325 (define_G),
:328-341 (define_F), :392-442 (define_D), :448-542 (GANLoss), :585-594 (Normalize),
:667-719 (PatchSampleF), :1051-1154 (ResnetGenerator), :1267-1324 (ResnetBlock),
:1649-1750 (Multiscale / NLayer discriminators).
"""
import math

import numpy as np
import torch
import torch.nn as nn
from torch.nn import init
from torch.optim import lr_scheduler

from . import ops
from .ops import (ACT_LRELU, ACT_NONE, ACT_RELU, FMT_BF16X2, FMT_F32, NORM_BATCH, NORM_INSTANCE, NORM_NONE,
                  PAD_REFLECT, PAD_ZERO)

TC_ENABLED = True  # tcgen05 path for eligible layers (64-multiple channels, stride 1)
PARALLEL_SCALES = True  # run the scales of a multiscale discriminator on parallel streams


def _require_cuda(t, who):
    if not t.is_cuda:
        raise RuntimeError("%s runs only on a CUDA device through libskit_b200.so; there is no CPU fallback "
                           "(use the reference or oracle/ for CPU checks)" % who)


# ----------------------------------------------------------------------------- parameter holders
class Conv2d(nn.Module):
    """Parameter holder with nn.Conv2d's state_dict keys (weight [co,ci,k,k], bias [co]).
    Class name contains 'Conv' so init_weights treats it like the reference's conv layers."""

    def __init__(self, ci, co, k, stride=1, bias=True):
        super().__init__()
        self.ci, self.co, self.k, self.stride = ci, co, k, stride
        self.allow_tc = True   # the U-Net generator keeps its (mostly thin) layers on the fp32 CUDA-core kernels
        self.tc_thin = False   # stride-1 layer with a thin side (9-channel stem, 5-channel head) routed to the halo tcgen05 kernel
        self.fold_in_cp = 0    # > 0: forward / wgrad run on the x-folded input operand (k*cp <= 64 channels, filter k x 1)
        self.fold_out_cp = 0   # > 0: the input gradient runs on the x-folded output-gradient operand
        self.tc_dgrad_s2 = False   # stride-2 layer with a thin input (PatchGAN stem 4/7 -> 64): only its INPUT GRADIENT runs on
                                   # tcgen05 (four parity sub-convolutions, K = co = 64, N = ci), forward and wgrad stay fp32
        self.weight = nn.Parameter(torch.empty(co, ci, k, k))
        self.bias = nn.Parameter(torch.zeros(co)) if bias else None
        self.reset_parameters()
        self._packs = {}

    def reset_parameters(self):
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        if self.bias is not None:
            bound = 1 / math.sqrt(self.ci * self.k * self.k)
            init.uniform_(self.bias, -bound, bound)

    @property
    def use_tc(self):
        if not (TC_ENABLED and self.allow_tc):
            return False
        if self.tc_thin and self.stride == 1:
            return True
        return self.stride in (1, 2) and self.ci % 64 == 0 and self.co % 64 == 0 and (self.stride == 1 or self.k % 2 == 0)

    @property
    def ci_pad(self):
        """Input channels of the tensor-core operand (zero-padded to a multiple of 16 for thin inputs)."""
        return self.ci if self.ci % 64 == 0 else -(-self.ci // 16) * 16

    @property
    def co_pad(self):
        """Channels of the tensor-core output-gradient operand (zero-padded to a multiple of 8 for thin outputs)."""
        return self.co if self.co % 64 == 0 else -(-self.co // 8) * 8

    def pack(self, mode):
        """mode 0 forward, 1 stride-1 dgrad (tensor-core), 2 gather dgrad (CUDA-core), 3 stride-2 phase dgrad (tensor-core)."""
        pk = self._packs.get(mode)
        if pk is None:
            bf16 = (self.use_tc and mode in (0, 1, 3)) or (self.tc_dgrad_s2 and mode == 3)
            kpad = 0
            if bf16 and self.tc_thin:
                kpad = self.ci_pad if mode == 0 else self.co_pad
            if bf16 and mode == 0 and self.fold_in_cp:
                pk = ops.PackedWeights(self.weight, 4, want_f32=False, want_bf16=True, cp=self.fold_in_cp)
            elif bf16 and mode == 1 and self.fold_out_cp:
                pk = ops.PackedWeights(self.weight, 5, want_f32=False, want_bf16=True, cp=self.fold_out_cp)
            else:
                pk = ops.PackedWeights(self.weight, mode, want_f32=not bf16, want_bf16=bf16, kpad=kpad)
            self._packs[mode] = pk
        return pk

    def refresh_packs(self):
        for pk in self._packs.values():
            pk.refresh(self.weight)

    def drop_packs(self):
        self._packs = {}


class BatchNorm2d(nn.Module):
    """Parameter/buffer holder with nn.BatchNorm2d's keys (networks.py:127-145 norm_layer)."""

    def __init__(self, c, momentum=0.1, eps=1e-5):
        super().__init__()
        self.c, self.momentum, self.eps = c, momentum, eps
        self.weight = nn.Parameter(torch.ones(c))
        self.bias = nn.Parameter(torch.zeros(c))
        self.register_buffer("running_mean", torch.zeros(c))
        self.register_buffer("running_var", torch.ones(c))
        self.register_buffer("num_batches_tracked", torch.tensor(0, dtype=torch.long))


class _Placeholder(nn.Module):
    """Keeps nn.Sequential indices aligned with the reference (ReLU, pads, norms without state)."""

    def __init__(self, what):
        super().__init__()
        self.what = what

    def extra_repr(self):
        return self.what


class _Blur(nn.Module):
    """Downsample / Upsample filter buffer `filt` (networks.py:51-107); the kernels hard-code the
    same [1,2,1] / [1,3,3,1] taps, the buffer exists for checkpoint compatibility."""

    def __init__(self, channels, filt_size, up):
        super().__init__()
        a = np.array([1.0, 2.0, 1.0] if filt_size == 3 else [1.0, 3.0, 3.0, 1.0])
        f = torch.tensor(a[:, None] * a[None, :], dtype=torch.float32)
        f = f / f.sum() * (4.0 if up else 1.0)
        self.register_buffer("filt", f[None, None].repeat(channels, 1, 1, 1))


class _FlatParamsMixin:
    """All parameters of a net live in one flat fp32 bucket (and one flat gradient bucket): the Adam
    kernel and the data-parallel all-reduce each touch exactly one tensor."""

    def flatten_parameters(self):
        params = [p for p in self.parameters()]
        dev = params[0].device
        total = sum(p.numel() for p in params)
        flat = torch.empty(total, dtype=torch.float32, device=dev)
        grad = torch.zeros(total, dtype=torch.float32, device=dev)
        off = 0
        for p in params:
            n = p.numel()
            flat[off:off + n].copy_(p.data.reshape(-1))
            p.data = flat[off:off + n].view(p.shape)
            p.grad = grad[off:off + n].view(p.shape)
            off += n
        self.flat_param, self.flat_grad = flat, grad
        self.exp_avg = torch.zeros_like(flat)
        self.exp_avg_sq = torch.zeros_like(flat)
        for m in self.modules():
            if hasattr(m, "drop_packs"):
                m.drop_packs()
        return self

    def ensure_flat(self):
        p = next(self.parameters())
        if getattr(self, "flat_param", None) is None or self.flat_param.device != p.device or \
                p.data.untyped_storage().data_ptr() != self.flat_param.untyped_storage().data_ptr():
            self.flatten_parameters()

    def zero_grad(self, set_to_none=False):  # grads are views of the flat bucket; never set to None
        if getattr(self, "flat_grad", None) is not None:
            if self.flat_grad.is_cuda:
                ops.zero_(self.flat_grad)       # a memset node in the captured step graph
            else:
                self.flat_grad.zero_()
        else:
            super().zero_grad(set_to_none=False)

    def refresh_packs(self):
        """Re-pack every weight of the net (after an optimiser step): one batched launch over a device table of all
        existing packs; the table is rebuilt (host -> device, outside any graph capture) only when a new pack appeared."""
        pairs = []
        for m in self.modules():
            if m is not self and hasattr(m, "_packs"):
                w = m.weight if m.weight.dim() == 4 else m.weight.view(m.weight.shape[0], m.weight.shape[1], 1, 1)
                for mode in sorted(m._packs):
                    pairs.append((w, m._packs[mode]))
        if not pairs:
            return
        key = tuple((w.data_ptr(), id(pk)) for w, pk in pairs)
        tab = self.__dict__.get("_pack_table")
        if tab is None or tab.key != key:
            if torch.cuda.is_current_stream_capturing():
                for w, pk in pairs:     # a pack created during capture: fall back to per-pack launches for this capture
                    pk.refresh(w)
                return
            tab = ops.PackTable(pairs, pairs[0][0].device)
            self.__dict__["_pack_table"] = tab
        tab.refresh()


# ----------------------------------------------------------------------------- weight gradients off the critical path
ASYNC_WGRAD = True
_WG = {}   # id of the launching stream -> (side stream, tensors kept alive until the join)


def _wgrad_async(fn, keep):
    """Nothing consumes a weight gradient before the optimiser, while the input-gradient chain is strictly serial:
    launch the wgrad kernels on a side stream (a parallel branch of the captured graph) so they fill the SMs the
    dgrad / elementwise kernels leave idle.  `keep`: the operands the side stream reads — held until `_wgrad_join`
    so that the caching allocator cannot hand their memory to a later main-stream allocation."""
    if not ASYNC_WGRAD:
        fn()
        return
    cur = torch.cuda.current_stream()
    ent = _WG.get(cur.cuda_stream)
    if ent is None:
        if len(_WG) > 64:      # every graph capture runs on a fresh stream: drop idle entries of streams long gone
            for k in [k for k, v in _WG.items() if not v[2]]:
                del _WG[k]
        ent = [torch.cuda.Stream(device=cur.device), [], False]
        _WG[cur.cuda_stream] = ent
    side, ka = ent[0], ent[1]
    side.wait_stream(cur)
    with torch.cuda.stream(side):
        fn()
    ka.extend(keep)
    ent[2] = True      # work is pending on the side stream since the last join


def _wgrad_side_stream():
    """The side stream carrying weight-gradient kernels launched from the current stream since its last join (or None)."""
    ent = _WG.get(torch.cuda.current_stream().cuda_stream)
    return ent[0] if ent is not None and ent[2] else None


def _wgrad_join():
    """Join only when something was launched since the last join: stream handles are recycled by the driver, so an entry may
    belong to a stream that no longer exists — waiting on its (idle) side stream from inside a graph capture would be an
    illegal dependency on uncaptured work."""
    cur = torch.cuda.current_stream()
    ent = _WG.get(cur.cuda_stream)
    if ent is not None and ent[2]:
        cur.wait_stream(ent[0])
        ent[1].clear()
        ent[2] = False


# ----------------------------------------------------------------------------- shared conv stage helpers
def _conv_fwd(layer, x_op, org, ho, wo, stats_mode, with_bias=True):
    pk = layer.pack(0)
    return ops.conv2d_fwd(x_op, pk, layer.stride, org, ho, wo, bias=layer.bias if with_bias else None, stats_mode=stats_mode)


def _stage_bwd(layer, x_op, raw, mr, norm_mode, act, count, dpad=None, pad=0, pad_mode=PAD_ZERO, dadd=None,
               gamma=None, beta=None, dgamma=None, dbeta=None, need_dgrad=True, need_wgrad=True, draw_add=None, dsum_out=None,
               dgrad_pack=None):
    """Backward of [conv -> norm -> act] given the gradient w.r.t. the activated output, either as
    a halo'd tensor `dpad` (gradient of the next conv's operand) and/or a dense `dadd`.
    Returns the gradient w.r.t. the conv's own haloed operand (NHWC fp32 [n, hp, wp, ci]) or None."""
    n, ho, wo, co = raw.shape
    if dpad is None and dadd is None:   # only a tapped-feature gradient on the raw output reaches this stage
        dadd = ops.zeros_big(raw.shape, torch.float32, raw.device)
    if dsum_out is not None:   # also materialise the incoming gradient fold(dpad) + dadd (a ResnetBlock's skip path needs it)
        g, sums, ds = ops.act_norm_bwd_reduce_ex(raw.shape, dpad=dpad, pad=pad, pad_mode=pad_mode, dadd=dadd, raw=raw, mr=mr,
                                                 norm_mode=norm_mode, act=act, gamma=gamma, beta=beta, want_dsum=True)
        dsum_out.append(ds)
    else:
        g, sums = ops.act_norm_bwd_reduce(raw.shape, dpad, pad, pad_mode, dadd, raw, mr, norm_mode, gamma, beta, act)
    tc = layer.use_tc
    s2_only = (not tc) and need_dgrad and layer.tc_dgrad_s2 and TC_ENABLED and co % 64 == 0
    q = 0 if (not (tc or s2_only) or not need_dgrad) else (layer.k - 1 if layer.stride == 1 else layer.k // 2 - 1)
    d_op = ops.norm_bwd_apply(g, raw, mr, norm_mode, gamma, sums, count, dgamma, dbeta, pad=q,
                              fmt=FMT_BF16X2 if (tc or s2_only) else FMT_F32, extra=draw_add)
    if need_wgrad:
        # A conv bias that feeds a norm layer has an identically zero gradient (the norm removes any constant
        # per-channel shift; d_raw sums to zero over the normalised axes) — the reference only accumulates
        # rounding noise there — so its reduction is skipped and the zeroed flat-grad entry stands.
        want_db = layer.bias is not None and (norm_mode == NORM_NONE or draw_add is not None)   # a raw-output tap sees the bias
        if tc and layer.fold_in_cp:
            def wg():
                ops.conv2d_wgrad_folded(x_op, d_op, q, layer.k, layer.k, layer.fold_in_cp, ho, wo, layer.weight.grad)
                if want_db:
                    ops.dbias(d_op, q, ho, wo, layer.bias.grad)
            _wgrad_async(wg, (x_op, d_op))
        else:
            _wgrad_async(lambda: ops.conv2d_wgrad(x_op, 0, d_op, q, layer.k, layer.stride, ho, wo, layer.weight.grad,
                                                  layer.bias.grad if want_db else None), (x_op, d_op))
    if not need_dgrad:
        return None
    if dgrad_pack is not None:   # input gradient for a subset of the input channels only (thin mode-1 pack of a weight slice)
        wp_in = x_op.wp + (layer.k - 1 if layer.fold_in_cp else 0)
        return ops.conv2d_fwd(d_op, dgrad_pack, 1, 0, x_op.hp, wp_in)[0]
    if (tc or s2_only) and layer.stride == 2:
        return ops.conv2d_dgrad_s2(d_op, q, layer.pack(3), layer.k, ho, wo, x_op.hp, x_op.wp)
    if tc:
        if layer.fold_out_cp == 0 and layer.pack(1).kw == 0:
            return ops.conv2d_dgrad_s1(d_op, layer.pack(1))     # dx has the size of the conv's own (haloed) input operand
        wp_in = x_op.wp + (layer.k - 1 if layer.fold_in_cp else 0)     # a folded operand is k-1 columns narrower than the input
        dx, _ = ops.conv2d_fwd(d_op, layer.pack(1), 1, 0, x_op.hp, wp_in)
        return dx
    return ops.conv2d_dgrad_gather(d_op.data, layer.pack(2), layer.stride, x_op.hp, x_op.wp)


# ----------------------------------------------------------------------------- generator
class ResnetBlock(nn.Module):
    """conv_block = [ReflectionPad2d(1), Conv2d, IN, ReLU, ReflectionPad2d(1), Conv2d, IN]
    (networks.py:1281-1320); out = x + conv_block(x) (:1322)."""

    def __init__(self, dim, use_bias=True):
        super().__init__()
        self.conv_block = nn.Sequential(
            _Placeholder("ReflectionPad2d(1)"), Conv2d(dim, dim, 3, bias=use_bias), _Placeholder("InstanceNorm2d"),
            _Placeholder("ReLU"), _Placeholder("ReflectionPad2d(1)"), Conv2d(dim, dim, 3, bias=use_bias),
            _Placeholder("InstanceNorm2d"))


class ResnetGenerator(_FlatParamsMixin, nn.Module):
    """Resnet generator with antialiased down/up-sampling (networks.py:1051-1154), InstanceNorm
    (affine=False), reflect padding, 5 output channels.  `model` index map: SURVEY.md A.2."""

    def __init__(self, input_nc, output_nc, ngf=64, n_blocks=9, opt=None, **unused):
        super().__init__()
        assert n_blocks >= 0
        self.input_nc, self.output_nc, self.ngf, self.n_blocks, self.opt = input_nc, output_nc, ngf, n_blocks, opt
        P = _Placeholder
        seq = [P("ReflectionPad2d(3)"), Conv2d(input_nc, ngf, 7), P("InstanceNorm2d"), P("ReLU")]
        for i in range(2):
            m = 2 ** i
            seq += [Conv2d(ngf * m, ngf * m * 2, 3), P("InstanceNorm2d"), P("ReLU"), _Blur(ngf * m * 2, 3, up=False)]
        seq += [ResnetBlock(ngf * 4) for _ in range(n_blocks)]
        for i in range(2):
            m = 2 ** (2 - i)
            seq += [_Blur(ngf * m, 4, up=True), Conv2d(ngf * m, ngf * m // 2, 3), P("InstanceNorm2d"), P("ReLU")]
        seq += [P("ReflectionPad2d(3)"), Conv2d(ngf, output_nc, 7), P("Tanh")]
        self.model = nn.Sequential(*seq)
        nb = n_blocks
        # unregistered aliases (object.__setattr__ keeps them out of state_dict)
        for name, idx in (("_c1", 1), ("_c4", 4), ("_c8", 8), ("_u1", 12 + nb + 1), ("_u2", 12 + nb + 5), ("_out", 12 + nb + 9)):
            object.__setattr__(self, name, self.model[idx])
        object.__setattr__(self, "_blocks", [self.model[12 + b] for b in range(nb)])
        if ngf % 64 == 0:   # tensor-core configuration: the thin 7x7 stem / head convs join the tcgen05 path (channel-padded)
            self._c1.tc_thin = True
            self._out.tc_thin = True
            if 7 * input_nc <= 64:      # 7x7 stem: fold the filter columns into the channel axis (7 x 64-channel K steps, not 49 thin ones)
                self._c1.fold_in_cp = input_nc
            if output_nc <= 8:          # 7x7 head: the same for its input gradient
                self._out.fold_out_cp = 8

    # -- explicit forward.  srcs: list of NCHW fp32 tensors whose channel concat is the input.
    def fwd(self, srcs, mask=None, scale_nz=0.25, save=True, want_normal=True, taps=None, style_code=None):
        _require_cuda(srcs[0], "ResnetGenerator")   # style_code: accepted and ignored like the reference (networks.py:1131)
        IN = NORM_INSTANCE
        n, _, S_h, S_w = srcs[0].shape
        ctx = {}
        feats = {}

        def tap(i, fn):
            if taps is not None and i in taps:
                feats[i] = fn()

        op0, thin0 = self._stem_operand(srcs)
        tap(0, lambda: thin0.data if thin0 is not None else (ops.nchw_cat_to_operand(srcs, 3, PAD_REFLECT).data if self._c1.use_tc else op0.data))
        # inference (save=False) drops every intermediate as soon as its consumer is enqueued: the forward then holds ~0.8 GB per
        # 1024x1024 image instead of ~5 GB, which is what bounds the batch size of the configs[4] sweep
        raw1, st = _conv_fwd(self._c1, op0, 0, S_h, S_w, IN)
        tap(1, lambda: raw1)
        _, op1, mr1 = ops.norm_act_pad_stats(raw1, st, S_h * S_w, IN, act=ACT_RELU, pad=1, pad_mode=PAD_ZERO, fmt=self._fmt(self._c4))
        if not save:
            del raw1, op0, thin0
        raw4, st = _conv_fwd(self._c4, op1, 0, S_h, S_w, IN)
        tap(4, lambda: raw4)
        a4, _, mr4 = ops.norm_act_pad_stats(raw4, st, S_h * S_w, IN, act=ACT_RELU, want_dense=True)
        if not save:
            del raw4, op1
        d4 = ops.blur_down_fwd(a4)
        del a4
        _, op4 = ops.norm_act_pad(d4, pad=1, pad_mode=PAD_ZERO, fmt=self._fmt(self._c8))
        del d4
        h2, w2 = S_h // 2, S_w // 2
        raw8, st = _conv_fwd(self._c8, op4, 0, h2, w2, IN)
        tap(8, lambda: raw8)
        a8, _, mr8 = ops.norm_act_pad_stats(raw8, st, h2 * w2, IN, act=ACT_RELU, want_dense=True)
        if not save:
            del raw8, op4
        t = ops.blur_down_fwd(a8)
        del a8
        h4, w4 = h2 // 2, w2 // 2
        blocks = []
        nb = self.n_blocks
        fmt_b = self._fmt(self._blocks[0].conv_block[1]) if nb else FMT_F32
        _, op_t = ops.norm_act_pad(t, pad=1, pad_mode=PAD_REFLECT, fmt=fmt_b) if nb else (None, None)
        for b, blk in enumerate(self._blocks):
            ca, cb = blk.conv_block[1], blk.conv_block[5]
            rawA, st = _conv_fwd(ca, op_t, 0, h4, w4, IN)
            _, opA, mrA = ops.norm_act_pad_stats(rawA, st, h4 * w4, IN, act=ACT_RELU, pad=1, pad_mode=PAD_REFLECT, fmt=self._fmt(cb))
            rawB, st = _conv_fwd(cb, opA, 0, h4, w4, IN)
            last = b == nb - 1
            t_new, op_next, mrB = ops.norm_act_pad_stats(rawB, st, h4 * w4, IN, residual=t, want_dense=True, pad=1, pad_mode=PAD_REFLECT,
                                                         fmt=None if last else fmt_b)
            if save:
                blocks.append((op_t, rawA, mrA, opA, rawB, mrB))
            t, op_t = t_new, op_next
            del rawA, opA, rawB, t_new, op_next
            tap(12 + b, lambda: t)
        u1 = ops.blur_up_fwd(t)
        _, op_u1 = ops.norm_act_pad(u1, pad=1, pad_mode=PAD_ZERO, fmt=self._fmt(self._u1))
        del u1
        if not save:
            del t, op_t
        raw22, st = _conv_fwd(self._u1, op_u1, 0, h2, w2, IN)
        a22, _, mr22 = ops.norm_act_pad_stats(raw22, st, h2 * w2, IN, act=ACT_RELU, want_dense=True)
        if not save:
            del raw22, op_u1
        u2 = ops.blur_up_fwd(a22)
        del a22
        _, op_u2 = ops.norm_act_pad(u2, pad=1, pad_mode=PAD_ZERO, fmt=self._fmt(self._u2))
        del u2
        raw26, st = _conv_fwd(self._u2, op_u2, 0, S_h, S_w, IN)
        _, op26, mr26 = ops.norm_act_pad_stats(raw26, st, S_h * S_w, IN, act=ACT_RELU, pad=3, pad_mode=PAD_REFLECT, fmt=self._fmt(self._out))
        if not save:
            del raw26, op_u2
        raw30, _ = _conv_fwd(self._out, op26, 0, S_h, S_w, NORM_NONE)
        fI, fT, fN = ops.g_head_fwd(raw30, mask, scale_nz, want_normal)
        if save:
            ctx.update(op0=op0, raw1=raw1, mr1=mr1, op1=op1, raw4=raw4, mr4=mr4, op4=op4, raw8=raw8, mr8=mr8,
                       blocks=blocks, op_u1=op_u1, raw22=raw22, mr22=mr22, op_u2=op_u2, raw26=raw26, mr26=mr26,
                       op26=op26, raw30=raw30, mask=mask, dims=(n, S_h, S_w))
        return (fI, fT, fN), ctx, feats

    @staticmethod
    def _fmt(layer):
        return FMT_BF16X2 if layer.use_tc else FMT_F32

    def _stem_operand(self, srcs):
        """Operand of the 7x7 stem conv: reflect-padded channel concat of the NCHW sources -> (operand, thin fp32 operand or
        None).  Tensor-core path: x-folded bf16x2 (fold_in_cp) or channel-padded bf16x2; otherwise plain fp32."""
        c1 = self._c1
        if c1.use_tc and c1.fold_in_cp:
            thin = ops.nchw_cat_to_operand(srcs, 3, PAD_REFLECT)
            return ops.fold_x(thin, c1.k), thin
        if c1.use_tc:
            return ops.nchw_cat_to_operand(srcs, 3, PAD_REFLECT, fmt=FMT_BF16X2, cpad=c1.ci_pad), None
        return ops.nchw_cat_to_operand(srcs, 3, PAD_REFLECT), None

    def tappable_layers(self):
        """nn.Sequential indices whose output the fused path materialises: padded input, the three
        stem convs (pre-norm) and every ResnetBlock output (CUT's nce_layers 0,4,8,12,16 are included)."""
        return {0, 1, 4, 8} | {12 + b for b in range(self.n_blocks)}

    # -- explicit backward: dI [n,3,h,w], dT [n,2,h,w] are gradients w.r.t. fake_I / fake_T (after *M).
    def grad_split_offset(self, block):
        """Offset (in elements of the flat buckets) of ResnetBlock `block`'s first parameter: everything from there on — the
        later blocks, the up-convs and the head — has its final gradient once the backward pass has finished that block."""
        first = next(self._blocks[block].parameters())
        return (first.data_ptr() - self.flat_param.data_ptr()) // 4

    def bwd(self, ctx, dI, dT, on_tail_done=None, tail_block=None):
        """on_tail_done(offset): called once, right after ResnetBlock `tail_block` (and everything behind it) has been
        processed: flat_grad[offset:] is final as soon as the weight-gradient side stream drains — the data-parallel step
        starts that bucket's all-reduce there, beside the rest of the backward pass."""
        IN = NORM_INSTANCE
        n, S_h, S_w = ctx["dims"]
        h2, w2, h4, w4 = S_h // 2, S_w // 2, S_h // 4, S_w // 4
        lo = self._out
        if lo.use_tc:   # 5-channel gradient, zero-padded to 8 channels and haloed by k-1 for the flipped-filter input gradient
            q = lo.k - 1
            d30 = ops.g_head_bwd(ctx["raw30"], ctx["mask"], dI, dT, q, fmt=FMT_BF16X2, cpad=lo.co_pad)
            op26 = ctx["op26"]
            d30f = ops.fold_x(d30, lo.k) if lo.fold_out_cp else d30
            if lo.fold_out_cp and op26.c % 64 == 0:     # 7 row taps against the folded gradient instead of 49 passes over op26
                def wg():
                    ops.conv2d_wgrad_dyfolded(op26, d30f, lo.k, lo.fold_out_cp, S_h, S_w, lo.weight.grad)
                    ops.dbias_n(d30, q, S_h, S_w, lo.co, lo.bias.grad)
                _wgrad_async(wg, (op26, d30, d30f))
            else:
                _wgrad_async(lambda: ops.conv2d_wgrad(op26, 0, d30, q, lo.k, 1, S_h, S_w, lo.weight.grad, lo.bias.grad), (op26, d30))
            dpad26, _ = ops.conv2d_fwd(d30f, lo.pack(1), 1, 0, S_h + 6, S_w + 6)
        else:
            d30 = ops.g_head_bwd(ctx["raw30"], ctx["mask"], dI, dT, 0)
            ops.conv2d_wgrad(ctx["op26"], 0, d30, 0, lo.k, 1, S_h, S_w, lo.weight.grad, lo.bias.grad)
            dpad26 = ops.conv2d_dgrad_gather(d30.data, lo.pack(2), 1, S_h + 6, S_w + 6)
        dpad_u2 = _stage_bwd(self._u2, ctx["op_u2"], ctx["raw26"], ctx["mr26"], IN, ACT_RELU, S_h * S_w,
                             dpad=dpad26, pad=3, pad_mode=PAD_REFLECT)
        du2, _ = ops.act_norm_bwd_reduce((n, S_h, S_w, dpad_u2.shape[3]), dpad=dpad_u2, pad=1, pad_mode=PAD_ZERO)
        da22 = ops.blur_up_bwd(du2)
        dpad_u1 = _stage_bwd(self._u1, ctx["op_u1"], ctx["raw22"], ctx["mr22"], IN, ACT_RELU, h2 * w2, dadd=da22)
        du1, _ = ops.act_norm_bwd_reduce((n, h2, w2, dpad_u1.shape[3]), dpad=dpad_u1, pad=1, pad_mode=PAD_ZERO)
        dt = ops.blur_up_bwd(du1)
        # per block: the output gradient dt feeds the conv branch and the skip; the next (earlier) block's output gradient is
        # fold(dpad_t) + dt, produced as a side output of that block's first norm-backward pass instead of a pass of its own
        pend = None   # dpad_t of the block processed before (the later block), not yet folded into dt
        for bi, blk, (op_t, rawA, mrA, opA, rawB, mrB) in zip(range(self.n_blocks - 1, -1, -1), reversed(self._blocks), reversed(ctx["blocks"])):
            ca, cb = blk.conv_block[1], blk.conv_block[5]
            if pend is None:
                dpadA = _stage_bwd(cb, opA, rawB, mrB, IN, ACT_NONE, h4 * w4, dadd=dt)
            else:
                out = []
                dpadA = _stage_bwd(cb, opA, rawB, mrB, IN, ACT_NONE, h4 * w4, dpad=pend, pad=1, pad_mode=PAD_REFLECT, dadd=dt, dsum_out=out)
                dt = out[0]
            pend = _stage_bwd(ca, op_t, rawA, mrA, IN, ACT_RELU, h4 * w4, dpad=dpadA, pad=1, pad_mode=PAD_REFLECT)
            if on_tail_done is not None and bi == tail_block:
                on_tail_done(self.grad_split_offset(bi))
        if pend is not None:
            dt, _ = ops.act_norm_bwd_reduce(dt.shape, dpad=pend, pad=1, pad_mode=PAD_REFLECT, dadd=dt)
        da8 = ops.blur_down_bwd(dt, h2, w2)
        dpad4 = _stage_bwd(self._c8, ctx["op4"], ctx["raw8"], ctx["mr8"], IN, ACT_RELU, h2 * w2, dadd=da8)
        dd4, _ = ops.act_norm_bwd_reduce((n, h2, w2, dpad4.shape[3]), dpad=dpad4, pad=1, pad_mode=PAD_ZERO)
        da4 = ops.blur_down_bwd(dd4, S_h, S_w)
        dpad1 = _stage_bwd(self._c4, ctx["op1"], ctx["raw4"], ctx["mr4"], IN, ACT_RELU, S_h * S_w, dadd=da4)
        _stage_bwd(self._c1, ctx["op0"], ctx["raw1"], ctx["mr1"], IN, ACT_RELU, S_h * S_w, dpad=dpad1, pad=1,
                   pad_mode=PAD_ZERO, need_dgrad=False)
        _wgrad_join()

    # -- encoder-only pass with saved activations (the PatchNCE query branch): features at `layers` (nn.Sequential
    #    indices from tappable_layers()), computed up to the deepest one only.
    def encode(self, srcs, layers):
        _require_cuda(srcs[0], "ResnetGenerator")
        IN = NORM_INSTANCE
        layers = sorted(set(int(i) for i in layers))
        bad = [i for i in layers if i not in self.tappable_layers()]
        if bad:
            raise NotImplementedError("feature taps %s are not exposed by the fused path (available: %s)" % (bad, sorted(self.tappable_layers())))
        top = layers[-1]
        n, _, S_h, S_w = srcs[0].shape
        feats, ctx = {}, dict(dims=(n, S_h, S_w), top=top)
        op0, thin0 = self._stem_operand(srcs)
        ctx["op0"] = op0
        if 0 in layers:
            feats[0] = thin0.data if thin0 is not None else (ops.nchw_cat_to_operand(srcs, 3, PAD_REFLECT).data if self._c1.use_tc else op0.data)
        if top == 0:
            return feats, ctx
        raw1, st = _conv_fwd(self._c1, op0, 0, S_h, S_w, IN)
        if 1 in layers:
            feats[1] = raw1
        if top == 1:
            ctx.update(raw1=raw1, mr1=ops.stats_finalize(st, S_h * S_w))
            return feats, ctx
        _, op1, mr1 = ops.norm_act_pad_stats(raw1, st, S_h * S_w, IN, act=ACT_RELU, pad=1, pad_mode=PAD_ZERO, fmt=self._fmt(self._c4))
        ctx.update(raw1=raw1, mr1=mr1)
        raw4, st = _conv_fwd(self._c4, op1, 0, S_h, S_w, IN)
        if 4 in layers:
            feats[4] = raw4
        if top == 4:
            ctx.update(op1=op1, raw4=raw4, mr4=ops.stats_finalize(st, S_h * S_w))
            return feats, ctx
        a4, _, mr4 = ops.norm_act_pad_stats(raw4, st, S_h * S_w, IN, act=ACT_RELU, want_dense=True)
        ctx.update(op1=op1, raw4=raw4, mr4=mr4)
        d4 = ops.blur_down_fwd(a4)
        del a4
        _, op4 = ops.norm_act_pad(d4, pad=1, pad_mode=PAD_ZERO, fmt=self._fmt(self._c8))
        h2, w2 = S_h // 2, S_w // 2
        raw8, st = _conv_fwd(self._c8, op4, 0, h2, w2, IN)
        if 8 in layers:
            feats[8] = raw8
        if top == 8:
            ctx.update(op4=op4, raw8=raw8, mr8=ops.stats_finalize(st, h2 * w2))
            return feats, ctx
        a8, _, mr8 = ops.norm_act_pad_stats(raw8, st, h2 * w2, IN, act=ACT_RELU, want_dense=True)
        ctx.update(op4=op4, raw8=raw8, mr8=mr8)
        t = ops.blur_down_fwd(a8)
        del a8
        h4, w4 = h2 // 2, w2 // 2
        nblk = top - 12 + 1
        fmt_b = self._fmt(self._blocks[0].conv_block[1])
        _, op_t = ops.norm_act_pad(t, pad=1, pad_mode=PAD_REFLECT, fmt=fmt_b)
        blocks = []
        for b in range(nblk):
            ca, cb = self._blocks[b].conv_block[1], self._blocks[b].conv_block[5]
            rawA, st = _conv_fwd(ca, op_t, 0, h4, w4, IN)
            _, opA, mrA = ops.norm_act_pad_stats(rawA, st, h4 * w4, IN, act=ACT_RELU, pad=1, pad_mode=PAD_REFLECT, fmt=self._fmt(cb))
            rawB, st = _conv_fwd(cb, opA, 0, h4, w4, IN)
            last = b == nblk - 1
            t, op_next, mrB = ops.norm_act_pad_stats(rawB, st, h4 * w4, IN, residual=t, want_dense=True, pad=1, pad_mode=PAD_REFLECT,
                                                     fmt=None if last else fmt_b)
            blocks.append((op_t, rawA, mrA, opA, rawB, mrB))
            op_t = op_next
            if 12 + b in layers:
                feats[12 + b] = t
        ctx["blocks"] = blocks
        return feats, ctx

    def encode_bwd(self, ctx, dfeats, input_channels=None):
        """dfeats: {layer: NHWC gradient of that tapped feature}.  Accumulates weight gradients and returns the gradient
        w.r.t. the reflect-padded input operand, NHWC [n, S+6, S+6, input_nc] (or None if nothing reaches it).
        input_channels = c: only the first c (<= 8) input channels' gradient is wanted (PatchNCE needs the sketch channel, not
        the positional encoding): on the tensor-core configuration the stem's input gradient then runs as a 7x7, 64 -> c
        convolution on the fp32 pipes (conv_head7_kernel) instead of an N = 16 MMA tile per 128 pixels, and a layer-0 tap
        gradient is NOT folded in (the caller adds its channels itself)."""
        IN = NORM_INSTANCE
        n, S_h, S_w = ctx["dims"]
        h2, w2, h4, w4 = S_h // 2, S_w // 2, S_h // 4, S_w // 4
        top = ctx["top"]
        dt = None
        if top >= 12:
            blocks = ctx["blocks"]
            for b in range(len(blocks) - 1, -1, -1):
                op_t, rawA, mrA, opA, rawB, mrB = blocks[b]
                ca, cb = self._blocks[b].conv_block[1], self._blocks[b].conv_block[5]
                tap = dfeats.get(12 + b)
                if dt is None:
                    dt = tap
                elif tap is not None:
                    dt, _ = ops.act_norm_bwd_reduce_ex(dt.shape, dadd=dt, dadd2=tap)
                if dt is None:
                    continue
                dpadA = _stage_bwd(cb, opA, rawB, mrB, IN, ACT_NONE, h4 * w4, dadd=dt)
                dpad_t = _stage_bwd(ca, op_t, rawA, mrA, IN, ACT_RELU, h4 * w4, dpad=dpadA, pad=1, pad_mode=PAD_REFLECT)
                dt, _ = ops.act_norm_bwd_reduce(dt.shape, dpad=dpad_t, pad=1, pad_mode=PAD_REFLECT, dadd=dt)
        dpad1 = None
        if top >= 8:
            da8 = ops.blur_down_bwd(dt, h2, w2) if dt is not None else None
            if da8 is not None or dfeats.get(8) is not None:
                dpad4 = _stage_bwd(self._c8, ctx["op4"], ctx["raw8"], ctx["mr8"], IN, ACT_RELU, h2 * w2, dadd=da8, draw_add=dfeats.get(8))
                dd4, _ = ops.act_norm_bwd_reduce((n, h2, w2, dpad4.shape[3]), dpad=dpad4, pad=1, pad_mode=PAD_ZERO)
            else:
                dd4 = None
        else:
            dd4 = None
        if top >= 4:
            da4 = ops.blur_down_bwd(dd4, S_h, S_w) if dd4 is not None else None
            if da4 is not None or dfeats.get(4) is not None:
                dpad1 = _stage_bwd(self._c4, ctx["op1"], ctx["raw4"], ctx["mr4"], IN, ACT_RELU, S_h * S_w, dadd=da4, draw_add=dfeats.get(4))
        dpad0 = None
        thin = input_channels is not None and self._c1.use_tc and 0 < input_channels <= 8
        if top >= 1 and (dpad1 is not None or dfeats.get(1) is not None):
            pk = None
            if thin:   # per-step pack of the weight slice [co, :c, 7, 7] (two tiny launches; the weights moved since the last step)
                pk = ops.PackedWeights(self._c1.weight.detach()[:, :input_channels].contiguous(), 1, want_f32=False, want_bf16=True)
            dpad0 = _stage_bwd(self._c1, ctx["op0"], ctx["raw1"], ctx["mr1"], IN, ACT_RELU, S_h * S_w, dpad=dpad1, pad=1,
                               pad_mode=PAD_ZERO, need_dgrad=True, draw_add=dfeats.get(1), dgrad_pack=pk)
        _wgrad_join()
        if thin:
            return dpad0
        if dfeats.get(0) is not None:
            if dpad0 is None:
                dpad0 = dfeats[0]
            else:
                dpad0, _ = ops.act_norm_bwd_reduce_ex(dpad0.shape, dadd=dpad0, dadd2=dfeats[0])
        return dpad0

    def feature_hw(self, layer, S_h, S_w):
        """Spatial size of a tapped feature (for drawing PatchSampleF ids on the host before the step)."""
        if layer == 0:
            return S_h + 6, S_w + 6
        if layer in (1, 4):
            return S_h, S_w
        if layer == 8:
            return S_h // 2, S_w // 2
        return S_h // 4, S_w // 4

    # -- reference module API (networks.py:1131-1154): returns tanh output [n, 5, h, w]
    def forward(self, input, layers=[], encode_only=False, style_code=None):
        _require_cuda(input, "ResnetGenerator")
        self.ensure_flat()
        self.refresh_packs()
        bad = [i for i in layers if i not in self.tappable_layers()]
        if bad:
            raise NotImplementedError("feature taps %s are not exposed by the fused path (available: %s)" % (bad, sorted(self.tappable_layers())))
        x = input.contiguous().float()
        (fI, fT, _), _, feats = self.fwd([x], mask=None, save=False, want_normal=False, taps=set(layers) if len(layers) else None)
        out = torch.cat([fI, fT], dim=1)
        if len(layers) > 0:
            fl = [feats[i].permute(0, 3, 1, 2) for i in layers if i in feats]  # NHWC storage, NCHW view
            return fl if encode_only else (out, fl)
        return out


# ----------------------------------------------------------------------------- default generator (U-Net)
class ConvTranspose2d(nn.Module):
    """Parameter holder with nn.ConvTranspose2d's state_dict keys: weight [ci][co][k][k], bias [co]
    (thirdparty/unet/unet_parts_custom.py:63).  Read as a conv weight with (out, in) = (ci, co) this is exactly
    the filter of the strided conv whose input gradient the transposed conv computes."""

    def __init__(self, ci, co, k=4, stride=2, padding=1):
        super().__init__()
        self.ci, self.co, self.k, self.stride, self.padding = ci, co, k, stride, padding
        self.weight = nn.Parameter(torch.empty(ci, co, k, k))
        self.bias = nn.Parameter(torch.zeros(co))
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        self._packs = {}

    def pack(self, mode):
        """mode 2: forward (gather form); mode 0: input gradient (strided conv of the haloed output gradient)."""
        pk = self._packs.get(mode)
        if pk is None:
            pk = ops.PackedWeights(self.weight, mode, want_f32=True, want_bf16=False)
            self._packs[mode] = pk
        return pk

    def refresh_packs(self):
        for pk in self._packs.values():
            pk.refresh(self.weight)

    def drop_packs(self):
        self._packs = {}


class _UnetDown(nn.Module):
    """Down (unet_parts_custom.py:9-37): [LeakyReLU(0.2, inplace)] -> Conv2d k4 s2 p1 -> [InstanceNorm2d]."""

    def __init__(self, ci, co, outermost=False, innermost=False):
        super().__init__()
        conv = Conv2d(ci, co, 4, stride=2)
        conv.allow_tc = False
        if outermost:
            seq = [conv]
        elif innermost:
            seq = [_Placeholder("LeakyReLU(0.2)"), conv]
        else:
            seq = [_Placeholder("LeakyReLU(0.2)"), conv, _Placeholder("InstanceNorm2d")]
        self.model = nn.Sequential(*seq)
        object.__setattr__(self, "conv", conv)


class _UnetUp(nn.Module):
    """Up (unet_parts_custom.py:40-80): ReLU(inplace) -> ConvTranspose2d k4 s2 p1 -> InstanceNorm2d | Tanh."""

    def __init__(self, ci, co, outermost=False):
        super().__init__()
        conv = ConvTranspose2d(ci, co)
        self.model = nn.Sequential(_Placeholder("ReLU"), conv, _Placeholder("Tanh" if outermost else "InstanceNorm2d"))
        object.__setattr__(self, "conv", conv)


class CustomUnetGenerator(_FlatParamsMixin, nn.Module):
    """`unet256_custom` (networks.py:1430-1645), the skitG / sinskitG default generator: 8 x Down, 8 x Up with twin
    RGB / touch decoders on the `num_layer_separate` outermost levels, InstanceNorm, 5 output channels.  Quirks kept:
    every skip tensor is the LeakyReLU'd activation (in-place aliasing), the outermost Up
    ignores its skip, skitG's style code is tiled and concatenated to the decoder input of the innermost
    `num_layer_style_code` levels.  All layers run on the fp32 CUDA-core kernels (channel counts 2..160)."""

    def __init__(self, input_nc, output_nc, num_downs=8, ngf=64, num_layer_separate=0, opt=None, input_size=1536, **unused):
        super().__init__()
        assert output_nc == 5, "current architecture is designed specifiacally for 5 output channels, 3 - RGB, 2 - touch"
        self.n_channels, self.output_nc, self.num_downs, self.opt, self.ngf = input_nc, output_nc, num_downs, opt, ngf
        o = vars(opt) if opt is not None else {}
        self.num_layer_style_code = num_downs if o.get("num_layer_style_code", -1) == -1 else o["num_layer_style_code"]
        use_style = bool(o.get("use_style_code", False))
        self.style_code_ncs = 0
        if use_style:
            if o.get("style_code_mode", "concat") != "concat" or o.get("style_code_mapping_mode", "tile") != "tile":
                raise NotImplementedError("B200 path: style code is built for style_code_mode='concat' with style_code_mapping_mode='tile' (the skitG default)")
            self.style_code_ncs = [o["style_code_dim"]] * self.num_layer_style_code
            for i in range(self.num_layer_style_code):   # unused in 'tile' mode; kept for checkpoint compatibility (A.1)
                out_sz = input_size // (2 ** (num_downs - i))
                setattr(self, "style_code_mapping%d" % i, nn.Sequential(
                    nn.Linear(o["style_code_dim"], out_sz * out_sz * o["style_code_dim"], bias=False),
                    _Placeholder("InstanceNorm1d"), _Placeholder("ReLU")))
        assert 0 <= num_layer_separate <= num_downs, "num_layer_separate should be in range [0, num_downs]"
        self.num_layer_separate = num_layer_separate
        out_main = output_nc - 2 if num_layer_separate > 0 else output_nc

        def style_nc(i):
            return self.style_code_ncs[num_downs - i - 1] if use_style and i >= num_downs - self.num_layer_style_code else 0

        def chans(i):   # (down in, down out) of level i
            if i == 0:
                return input_nc, ngf
            if i < num_downs // 2:
                return ngf * 2 ** (i - 1), ngf * 2 ** i
            return ngf * 8, ngf * 8

        self._style_nc = [style_nc(i) for i in range(num_downs)]
        for i in range(num_downs):
            ci, co = chans(i)
            setattr(self, "down%d" % i, _UnetDown(ci, co, outermost=i == 0, innermost=i == num_downs - 1))
            scaler = 1 if i in (0, num_downs - 1) else 2
            up_in = scaler * co + self._style_nc[i]
            up_out = (out_main if i == 0 else ci)
            setattr(self, "up%d" % i, _UnetUp(up_in, up_out, outermost=i == 0))
            if num_layer_separate >= i + 1:
                setattr(self, "up%d_T" % i, _UnetUp(up_in, 2 if i == 0 else ci, outermost=i == 0))

    def _up(self, i, T=False):
        m = getattr(self, "up%d%s" % (i, "_T" if T else ""), None)
        return None if m is None else m.conv

    def tappable_layers(self):
        return set()

    # -- explicit forward.  srcs: NCHW fp32 tensors whose channel concat is the input.
    def fwd(self, srcs, mask=None, scale_nz=0.25, save=True, want_normal=True, taps=None, style_code=None):
        _require_cuda(srcs[0], "CustomUnetGenerator")
        IN = NORM_INSTANCE
        nd = self.num_downs
        n, _, S_h, S_w = srcs[0].shape
        dev = srcs[0].device
        if S_h % (1 << nd) or S_w % (1 << nd):
            raise ValueError("CustomUnetGenerator: input %dx%d is not divisible by 2^%d" % (S_h, S_w, nd))
        op = ops.nchw_cat_to_operand(srcs, 1, PAD_ZERO)
        downs, h, w = [], S_h, S_w
        for i in range(nd):
            conv = getattr(self, "down%d" % i).conv
            h, w = h // 2, w // 2
            mode = IN if 0 < i < nd - 1 else NORM_NONE
            raw, st = ops.conv2d_fwd(op, conv.pack(0), 2, 0, h, w, bias=conv.bias, stats_mode=mode, impl=ops.IMPL_SIMT)
            mr = ops.stats_finalize(st, h * w) if mode != NORM_NONE else None
            downs.append((op, raw, mr, mode))
            if i < nd - 1:   # next level's input = LeakyReLU(norm(raw)), zero halo 1 — also the (aliased) skip tensor
                _, op = ops.norm_act_pad(raw, mr, mode, act=ACT_LRELU, pad=1, pad_mode=PAD_ZERO, fmt=FMT_F32)
        style = None
        if style_code is not None:
            style = torch.relu(style_code.to(dev, torch.float32))   # the decoder's ReLU acts on the concatenated input
        raw30 = torch.empty((n, S_h, S_w, 5), dtype=torch.float32, device=dev)
        # x / x_T: the current decoder activations as (raw, mean_rstd, norm_mode), normalised lazily on load
        x = (downs[nd - 1][1], None, NORM_NONE)
        xT = None
        ups = [None] * nd
        for i in range(nd - 1, -1, -1):
            has_skip = 0 < i < nd - 1
            cst = self._style_nc[i] if style is not None else 0
            if self._style_nc[i] and style is None:
                raise RuntimeError("CustomUnetGenerator was built with use_style_code but forward got style_code=None")

            def build(xx):
                craw = xx[0].shape[3]
                hh, ww = xx[0].shape[1:3]
                ctot = craw + cst + (downs[i][1].shape[3] if has_skip else 0)
                U = ops.Operand(n, hh, ww, ctot, 0, FMT_F32, dev)
                ops.norm_act_pad_into(U, 0, xx[0], xx[1], xx[2], act=ACT_RELU)
                if cst:
                    U.data[..., craw:craw + cst] = style[:, None, None, :]
                if has_skip:
                    ops.norm_act_pad_into(U, craw + cst, downs[i][1], downs[i][2], downs[i][3], act=ACT_RELU)
                return U, craw

            U, cx = build(x)
            convT = self._up(i, T=True)
            recT = None
            if convT is not None:
                shared = xT is None
                UT = U if shared else build(xT)[0]
                if i == 0:
                    ops.conv_transpose2d_fwd(UT.data, convT.pack(2), 2, 1, bias=convT.bias, out=raw30, out_c0=3)
                    xT = None
                else:
                    rT, st = ops.conv_transpose2d_fwd(UT.data, convT.pack(2), 2, 1, bias=convT.bias, stats_mode=IN)
                    xT = (rT, ops.stats_finalize(st, rT.shape[1] * rT.shape[2]), IN)
                recT = (UT, shared, xT)
            conv = self._up(i)
            if i == 0:
                ops.conv_transpose2d_fwd(U.data, conv.pack(2), 2, 1, bias=conv.bias, out=raw30, out_c0=0)
                x = None
            else:
                r, st = ops.conv_transpose2d_fwd(U.data, conv.pack(2), 2, 1, bias=conv.bias, stats_mode=IN)
                x = (r, ops.stats_finalize(st, r.shape[1] * r.shape[2]), IN)
            ups[i] = (U, cx, cst, x, recT)
        fI, fT, fN = ops.g_head_fwd(raw30, mask, scale_nz, want_normal)
        ctx = dict(downs=downs, ups=ups, raw30=raw30, mask=mask, dims=(n, S_h, S_w)) if save else {}
        return (fI, fT, fN), ctx, {}

    @staticmethod
    def _convT_bwd(layer, U, d_op, hin, win, want_dbias):
        """Backward of y = convT(U): d_op is dy haloed by 1 (fp32).  -> dU dense [n, hin, win, C_U]."""
        ops.conv2d_wgrad(d_op, 0, U, 0, layer.k, layer.stride, hin, win, layer.weight.grad, None, impl=ops.IMPL_SIMT)
        if want_dbias:
            ops.dbias(d_op, 1, 2 * hin, 2 * win, layer.bias.grad)
        dU, _ = ops.conv2d_fwd(d_op, layer.pack(0), layer.stride, 0, hin, win, impl=ops.IMPL_SIMT)
        return dU

    # -- explicit backward: dI [n,3,h,w], dT [n,2,h,w] are gradients w.r.t. fake_I / fake_T (after *M).
    def bwd(self, ctx, dI, dT):
        IN = NORM_INSTANCE
        nd = self.num_downs
        n, S_h, S_w = ctx["dims"]
        downs, ups = ctx["downs"], ctx["ups"]
        twin0 = self._up(0, T=True) is not None
        if twin0:
            d_main, d_T = ops.g_head_bwd_split(ctx["raw30"], ctx["mask"], dI, dT, 1)
        else:
            d_main, d_T = ops.g_head_bwd(ctx["raw30"], ctx["mask"], dI, dT, 1), None
        # decoder, outermost level first.  dU[i] = list of (dense gradient of level i's decoder input, owner) pairs
        dU = [None] * nd
        for i in range(nd):
            U, cx, cst, x_out, recT = ups[i]
            hin, win = U.h, U.w
            gm = self._convT_bwd(self._up(i), U, d_main, hin, win, want_dbias=i == 0)
            gT, shared = None, False
            if recT is not None:
                UT, shared, _ = recT
                gT = self._convT_bwd(self._up(i, T=True), UT, d_T, hin, win, want_dbias=i == 0)
            dU[i] = (gm, gT, shared)
            if i == nd - 1:
                break
            # gradients w.r.t. the raw outputs of level i+1's transposed convs (the x-part of U / U_T)
            nxt = ups[i + 1]
            ctot = U.c

            def raw_grad(xrec, g1, g2):
                raw, mr, mode = xrec
                g, sums = ops.act_norm_bwd_reduce_ex(raw.shape, dadd=g1, dadd2=g2, dadd_c0=0, dadd_ctot=ctot, raw=raw, mr=mr,
                                                     norm_mode=mode, act=ACT_RELU)
                return ops.norm_bwd_apply(g, raw, mr, mode, None, sums, raw.shape[1] * raw.shape[2], pad=1, fmt=FMT_F32)

            x_next, recT_next = nxt[3], nxt[4]
            if recT is not None and shared:        # first split level: one shared input, both decoders feed level i+1's main output
                d_main, d_T = raw_grad(x_next, gm, gT), None
            else:
                d_main = raw_grad(x_next, gm, None)
                d_T = raw_grad(recT_next[2], gT, None) if gT is not None else None
        # encoder, innermost level first
        dpad = None
        for i in range(nd - 1, -1, -1):
            op_in, raw, mr, mode = downs[i]
            conv = getattr(self, "down%d" % i).conv
            _, ho, wo, _ = raw.shape
            gm, gT, shared = dU[i]
            U, cx, cst = ups[i][0], ups[i][1], ups[i][2]
            if i == nd - 1:      # consumed by up_{nd-1} only, through ReLU (x-part of U)
                g, sums = ops.act_norm_bwd_reduce_ex(raw.shape, dadd=gm, dadd2=gT, dadd_c0=0, dadd_ctot=U.c, raw=raw, act=ACT_RELU)
            elif i == 0:         # consumed by down1 only (the outermost Up ignores its skip)
                g, sums = ops.act_norm_bwd_reduce_ex(raw.shape, dpad=dpad, pad=1, pad_mode=PAD_ZERO, raw=raw, act=ACT_LRELU)
            else:                # LeakyReLU'd tensor feeds down_{i+1} and, through the decoder's ReLU, the skip slice of U_i
                g, sums = ops.act_norm_bwd_reduce_ex(raw.shape, dpad=dpad, pad=1, pad_mode=PAD_ZERO, dadd=gm, dadd2=gT,
                                                     dadd_c0=cx + cst, dadd_ctot=U.c, dadd_relu_mask=True, raw=raw, mr=mr,
                                                     norm_mode=mode, act=ACT_LRELU)
            d_op = ops.norm_bwd_apply(g, raw, mr, mode, None, sums, ho * wo, pad=0, fmt=FMT_F32)
            ops.conv2d_wgrad(op_in, 0, d_op, 0, conv.k, 2, ho, wo, conv.weight.grad,
                             conv.bias.grad if mode == NORM_NONE else None, impl=ops.IMPL_SIMT)
            if i > 0:
                dpad = ops.conv2d_dgrad_gather(d_op.data, conv.pack(2), 2, op_in.hp, op_in.wp)

    # -- reference module API (networks.py:1538): returns the tanh output [n, 5, h, w]
    def forward(self, x, verbose=False, style_code=None):
        _require_cuda(x, "CustomUnetGenerator")
        self.ensure_flat()
        self.refresh_packs()
        (fI, fT, _), _, _ = self.fwd([x.contiguous().float()], mask=None, save=False, want_normal=False, style_code=style_code)
        return torch.cat([fI, fT], dim=1)


# ----------------------------------------------------------------------------- discriminators
class NLayerDiscriminator(_FlatParamsMixin, nn.Module):
    """PatchGAN (networks.py:1696-1750): k4, pad 2, strides 2,2,2,1,1; norm after convs 1..n_layers."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm="batch", **unused):
        super().__init__()
        self.n_layers, self.norm = n_layers, norm
        self.model = nn.Sequential(*self._build(input_nc, ndf, n_layers, norm))
        self._index_layers(self.model)

    @staticmethod
    def _build(input_nc, ndf, n_layers, norm):
        def nl(c):
            return BatchNorm2d(c) if norm == "batch" else _Placeholder("InstanceNorm2d" if norm == "instance" else "Identity")
        stem = Conv2d(input_nc, ndf, 4, stride=2)
        stem.tc_dgrad_s2 = ndf % 64 == 0
        seq = [stem, _Placeholder("LeakyReLU(0.2)")]
        nf = ndf
        for _ in range(1, n_layers):
            nf_prev, nf = nf, min(nf * 2, 512)
            seq += [Conv2d(nf_prev, nf, 4, stride=2), nl(nf), _Placeholder("LeakyReLU(0.2)")]
        nf_prev, nf = nf, min(nf * 2, 512)
        seq += [Conv2d(nf_prev, nf, 4, stride=1), nl(nf), _Placeholder("LeakyReLU(0.2)")]
        seq += [Conv2d(nf, 1, 4, stride=1)]
        return seq

    def _index_layers(self, seq):
        """[(conv, norm_module_or_None, act)] in execution order."""
        self._stages = _d_stages(seq, self.norm)

    def fwd(self, srcs, save=True, update_running=True):
        return _d_fwd(self._stages, self.norm, srcs, save, update_running)

    def bwd(self, ctx, dpred, need_wgrad=True, need_input_grad=False):
        return _d_bwd(self._stages, self.norm, ctx, dpred, need_wgrad, need_input_grad)

    def forward(self, input):
        _require_cuda(input, "NLayerDiscriminator")
        self.ensure_flat()
        self.refresh_packs()
        pred, _ = self.fwd([input.contiguous().float()], save=False, update_running=self.training)
        return pred.permute(0, 3, 1, 2)


def _d_stages(seq, norm):
    stages, mods, i = [], list(seq), 0
    while i < len(mods):
        conv = mods[i]
        assert isinstance(conv, Conv2d)
        nm, act = None, ACT_NONE
        j = i + 1
        has_norm = False
        while j < len(mods) and not isinstance(mods[j], Conv2d):
            if isinstance(mods[j], BatchNorm2d):
                nm, has_norm = mods[j], True
            elif isinstance(mods[j], _Placeholder) and mods[j].what == "InstanceNorm2d":
                has_norm = True
            elif isinstance(mods[j], _Placeholder) and mods[j].what.startswith("LeakyReLU"):
                act = ACT_LRELU
            j += 1
        mode = NORM_NONE if not has_norm else (NORM_BATCH if norm == "batch" else NORM_INSTANCE)
        stages.append((conv, nm, mode, act))
        i = j
    return stages


def _d_fwd(stages, norm, srcs, save, update_running=True, deferred=None):
    """srcs: NCHW tensors (channel concat = D input).  Returns (pred NHWC [n,h,w,1], ctx).
    deferred: a list — BatchNorm running-stat updates are appended to it as (stats, count, bn) instead of being applied,
    so that passes running on parallel streams can apply them afterwards in the reference's order
    (`apply_running_updates`)."""
    _require_cuda(srcs[0], "discriminator")
    n, _, h, w = srcs[0].shape
    x_op = ops.nchw_cat_to_operand(srcs, 2, PAD_ZERO)
    saved = []
    for si, (conv, bn, mode, act) in enumerate(stages):
        ho, wo = h // conv.stride + 1, w // conv.stride + 1   # floor((h + 4 - 4) / s) + 1
        raw, st = _conv_fwd(conv, x_op, 0, ho, wo, mode)
        mr = None
        if mode != NORM_NONE:
            cnt = n * ho * wo if mode == NORM_BATCH else ho * wo
            upd = bn is not None and update_running
            if upd and deferred is not None:
                deferred.append((st, cnt, bn))
                upd = False
            if upd:
                mr = ops.stats_finalize(st, cnt, bn.eps, bn.running_mean, bn.running_var, bn.momentum)
                bn.num_batches_tracked += 1
        if si == len(stages) - 1:
            if save:
                saved.append((x_op, raw, mr, (h, w)))
            pred = raw
            break
        nxt = stages[si + 1][0]
        fmt_n = FMT_BF16X2 if nxt.use_tc else FMT_F32
        gamma, beta = (bn.weight, bn.bias) if bn is not None else (None, None)
        x_in = x_op
        if mode != NORM_NONE and mr is None:     # statistics finalised inside the normalise / activate / pad pass
            _, x_op, mr = ops.norm_act_pad_stats(raw, st, cnt, mode, gamma, beta, act, pad=2, pad_mode=PAD_ZERO, fmt=fmt_n,
                                                 eps=bn.eps if bn is not None else 1e-5)
        else:
            _, x_op = ops.norm_act_pad(raw, mr, mode, gamma, beta, act, pad=2, pad_mode=PAD_ZERO, fmt=fmt_n)
        if save:
            saved.append((x_in, raw, mr, (h, w)))
        h, w = ho, wo
    return pred, dict(saved=saved, n=n)


def apply_running_updates(deferred):
    """Apply the BatchNorm running-stat updates collected by `_d_fwd(..., deferred=...)`, in list order."""
    for st, cnt, bn in deferred:
        ops.stats_finalize(st, cnt, bn.eps, bn.running_mean, bn.running_var, bn.momentum)
        bn.num_batches_tracked += 1
    deferred.clear()


def _d_bwd(stages, norm, ctx, dpred, need_wgrad=True, need_input_grad=False):
    """dpred: NHWC [n,h,w,1] gradient w.r.t. the prediction map.  Returns the gradient w.r.t. the
    haloed input operand (NHWC fp32 [n, H+4, W+4, cin]) if need_input_grad."""
    n = ctx["n"]
    dpad, dadd = None, dpred
    for si in range(len(stages) - 1, -1, -1):
        conv, bn, mode, act = stages[si]
        x_op, raw, mr, _ = ctx["saved"][si]
        _, ho, wo, _ = raw.shape
        cnt = n * ho * wo if mode == NORM_BATCH else ho * wo
        last_stage = si == len(stages) - 1
        dgrad = si > 0 or need_input_grad
        dx = _stage_bwd(conv, x_op, raw, mr, mode if not last_stage else NORM_NONE, act if not last_stage else ACT_NONE, cnt,
                        dpad=dpad, pad=2, pad_mode=PAD_ZERO, dadd=dadd,
                        gamma=bn.weight if bn is not None else None, beta=bn.bias if bn is not None else None,
                        dgamma=bn.weight.grad if (bn is not None and need_wgrad) else None,
                        dbeta=bn.bias.grad if (bn is not None and need_wgrad) else None,
                        need_dgrad=dgrad, need_wgrad=need_wgrad)
        dpad, dadd = dx, None
    _wgrad_join()
    return dpad


class MultiscaleDiscriminator(_FlatParamsMixin, nn.Module):
    """num_D PatchGANs over an AvgPool2d(3,2,1,count_include_pad=False) pyramid
    (networks.py:1649-1693): result[i] = [layer{num_D-1-i}(input downsampled i times)]."""

    def __init__(self, input_nc, ndf=64, n_layers=3, norm="batch", num_D=3, opt=None):
        super().__init__()
        self.n_layers, self.num_D, self.norm = n_layers, num_D, norm
        self.getIntermFeat = bool(getattr(opt, "getIntermFeat_D", False)) if opt is not None else False
        if self.getIntermFeat:
            raise NotImplementedError("getIntermFeat_D is not supported on the B200 path")
        for i in range(num_D):
            setattr(self, "layer%d" % i, nn.Sequential(*NLayerDiscriminator._build(input_nc, ndf, n_layers, norm)))
        self._scales = [_d_stages(getattr(self, "layer%d" % i), norm) for i in range(num_D)]

    def _side_streams(self, device):
        """One side stream per extra scale: the scales own disjoint parameters and buffers, and their (small, latency
        bound) kernels overlap when each scale runs on its own stream — also inside a captured CUDA graph, where
        the fork / join becomes parallel branches."""
        tab = getattr(self, "_streams", None)
        if tab is None:
            tab = {}
            object.__setattr__(self, "_streams", tab)
        key = (str(device), torch.cuda.current_stream().cuda_stream)   # passes launched from different streams do not share
        if key not in tab:
            tab[key] = [torch.cuda.Stream(device=device) for _ in range(self.num_D - 1)]
        return tab[key]

    def fwd(self, srcs, save=True, update_running=True, deferred=None):
        """-> (list of pred NHWC per scale i (full res first), ctx)"""
        pyramid = [srcs]
        for i in range(1, self.num_D):
            pyramid.append([ops.avgpool3s2_fwd(s) for s in pyramid[-1]])
        main = torch.cuda.current_stream()
        sides = self._side_streams(srcs[0].device) if PARALLEL_SCALES and self.num_D > 1 else []
        preds, ctxs = [None] * self.num_D, [None] * self.num_D
        for i in range(self.num_D):
            st = sides[i - 1] if (sides and i > 0) else main
            if st is not main:
                st.wait_stream(main)
            with torch.cuda.stream(st):
                dl = [] if deferred is not None else None
                preds[i], ctxs[i] = _d_fwd(self._scales[self.num_D - 1 - i], self.norm, pyramid[i], save, update_running, dl)
                if dl:
                    deferred.extend(dl)
        for st in sides:
            main.wait_stream(st)
        hw = [tuple(s.shape[-2:]) for s in srcs[:1]]
        return preds, dict(scales=ctxs, in_hw=hw[0], chans=[int(s.shape[1]) for s in srcs])

    def bwd(self, ctx, dpreds, need_wgrad=True, input_slice=None):
        """dpreds: per-scale NHWC gradients.  input_slice=(c0, cs): also return the NCHW gradient
        w.r.t. that channel slice of the (concatenated) input, summed over the pyramid."""
        H, W = ctx["in_hw"]
        sizes = [(H, W)]
        for _ in range(1, self.num_D):
            sizes.append(((sizes[-1][0] - 1) // 2 + 1, (sizes[-1][1] - 1) // 2 + 1))
        main = torch.cuda.current_stream()
        sides = self._side_streams(dpreds[0].device) if PARALLEL_SCALES and self.num_D > 1 else []
        dins = [None] * self.num_D
        for i in range(self.num_D):
            st = sides[i - 1] if (sides and i > 0) else main
            if st is not main:
                st.wait_stream(main)
            with torch.cuda.stream(st):
                dins[i] = _d_bwd(self._scales[self.num_D - 1 - i], self.norm, ctx["scales"][i], dpreds[i], need_wgrad,
                                 need_input_grad=input_slice is not None)
        for st in sides:
            main.wait_stream(st)
        if input_slice is None:
            return None
        c0, cs = input_slice
        acc = None
        for i in range(self.num_D - 1, -1, -1):
            h, w = sizes[i]
            gi = ops.operand_grad_to_nchw(dins[i], h, w, 2, PAD_ZERO, c0, cs)
            if acc is not None:
                ops.avgpool3s2_bwd(acc, h, w, dx=gi, accumulate=True)
            acc = gi
        return acc

    def forward(self, input):
        _require_cuda(input, "MultiscaleDiscriminator")
        self.ensure_flat()
        self.refresh_packs()
        preds, _ = self.fwd([input.contiguous().float()], save=False, update_running=self.training)
        return [[p.permute(0, 3, 1, 2)] for p in preds]


# ----------------------------------------------------------------------------- GAN loss
class GANLoss(nn.Module):
    """GANLoss (networks.py:448-542) in all of its modes on one fused kernel (value here; value + gradient inside the train step).
    'nonsaturating' / 'hinge' return the per-sample means [bs]; 'lsgan' / 'vanilla' / 'wgan' / 'wgangp' the mean over everything (a
    0-d tensor), as `nn.MSELoss` / `nn.BCEWithLogitsLoss` / `.mean()` do.  Multiscale predictions sum over scales; a bare tensor uses
    `input[-1]`, i.e. the LAST batch element only (reference quirk, :541-542)."""

    def __init__(self, gan_mode, target_real_label=1.0, target_fake_label=0.0):
        super().__init__()
        self.register_buffer("real_label", torch.tensor(target_real_label))
        self.register_buffer("fake_label", torch.tensor(target_fake_label))
        self.gan_mode = gan_mode
        if gan_mode not in ("lsgan", "vanilla", "wgan", "wgangp", "nonsaturating", "hinge"):
            raise NotImplementedError("gan mode %s not implemented" % gan_mode)
        self._labels = (float(target_real_label), float(target_fake_label))

    def get_loss_for_single_scale_discriminator(self, prediction, target_is_real):
        _require_cuda(prediction, "GANLoss")
        bs = prediction.shape[0]
        loss = torch.zeros(bs, dtype=torch.float32, device=prediction.device)
        ops.gan_loss(prediction.contiguous().float(), self.gan_mode, target_is_real, self._labels[0 if target_is_real else 1], loss)
        if self.gan_mode in ("nonsaturating", "hinge"):
            return loss
        return loss.mean()          # equal-sized samples: the mean of the per-sample means is the mean over all elements

    def __call__(self, input, target_is_real):
        if isinstance(input[0], list):
            loss = 0
            for input_i in input:
                loss = loss + self.get_loss_for_single_scale_discriminator(input_i[-1], target_is_real)
            return loss
        return self.get_loss_for_single_scale_discriminator(input[-1], target_is_real)


# ----------------------------------------------------------------------------- PatchNCE feature sampler
class Normalize(nn.Module):
    """x / (||x||_p + 1e-7) (networks.py:585-594); p = 2 runs inside the sampling kernel."""

    def __init__(self, power=2):
        super().__init__()
        self.power = power

    def forward(self, x):
        norm = x.pow(self.power).sum(1, keepdim=True).pow(1.0 / self.power)
        return x.div(norm + 1e-7)


class Linear(nn.Module):
    """Parameter holder with nn.Linear's keys (weight [out, in], bias [out]); runs as a 1x1 convolution over the
    sampled rows laid out as an NHWC map [1, rows, 1, in].  The class name keeps init_weights' 'Linear' dispatch."""

    def __init__(self, ci, co):
        super().__init__()
        self.ci, self.co = ci, co
        self.weight = nn.Parameter(torch.empty(co, ci))
        self.bias = nn.Parameter(torch.zeros(co))
        init.kaiming_uniform_(self.weight, a=math.sqrt(5))
        self._packs = {}

    def pack(self, mode):
        pk = self._packs.get(mode)
        if pk is None:
            pk = ops.PackedWeights(self.weight.view(self.co, self.ci, 1, 1), mode, want_f32=True, want_bf16=False)
            self._packs[mode] = pk
        return pk

    def refresh_packs(self):
        for pk in self._packs.values():
            pk.refresh(self.weight.view(self.co, self.ci, 1, 1))

    def drop_packs(self):
        self._packs = {}


def first_of_permutation(n, p):
    """The first `p` entries of a uniformly random permutation of range(n) — what PatchSampleF draws with
    `np.random.permutation(H * W)[:num_patches]` (networks.py:703-705) — in O(p) instead of O(n): Fisher-Yates from the front
    (slot i swaps with a uniform j in [i, n)), keeping only the touched slots in a dict.  Same distribution (p distinct indices,
    every ordered p-tuple equally likely), driven by the same legacy global NumPy stream (np.random.seed controls it); at
    774 x 774 positions the full permutation costs ~8 ms of host time per layer per step, this ~0.1 ms."""
    p = int(min(p, n))
    js = np.arange(p) + np.floor(np.random.random_sample(p) * (n - np.arange(p))).astype(np.int64)
    moved, out = {}, np.empty(p, dtype=np.int64)
    for i in range(p):
        j = int(js[i])
        vi, vj = moved.get(i, i), moved.get(j, j)
        out[i] = vj
        moved[j] = vi
    return out


class PatchSampleF(_FlatParamsMixin, nn.Module):
    """PatchSampleF (networks.py:667-719): gather `num_patches` spatial positions (shared across the batch) from each
    feature map, optionally run them through a per-layer 2-layer MLP created on first use (netF='mlp_sample'), and
    L2-normalise — one warp per sampled row for the gather / normalisation, 1x1-conv kernels for the MLP."""

    def __init__(self, use_mlp=False, init_type="normal", init_gain=0.02, nc=256, gpu_ids=[]):
        super().__init__()
        self.l2norm = Normalize(2)
        self.use_mlp, self.nc, self.mlp_init = use_mlp, nc, False
        self.init_type, self.init_gain, self.gpu_ids = init_type, init_gain, gpu_ids

    def create_mlp(self, feats=None, channels=None, device=None):
        """networks.py:678-686.  feats: NCHW feature list (reference signature) or `channels`: their channel counts."""
        if channels is None:
            channels = [int(f.shape[1]) for f in feats]
            device = feats[0].device
        for mlp_id, input_nc in enumerate(channels):
            setattr(self, "mlp_%d" % mlp_id, nn.Sequential(Linear(input_nc, self.nc), _Placeholder("ReLU"), Linear(self.nc, self.nc)))
        init_weights(self, self.init_type, init_gain=self.init_gain)
        if device is not None:
            self.to(device)
        self.mlp_init = True
        self.flat_param = None

    # -- explicit forward / backward on one NHWC feature map
    def sample_fwd(self, feat, ids, feat_id=0, save=False):
        """feat: NHWC [b, h, w, c]; ids: int32 device tensor [P] -> (normalised rows [b*P, nc or c], ctx)."""
        out0, rows = ops.patch_sample_l2norm(feat, ids, keep_pre=True)
        if not self.use_mlp:
            return out0, (dict(rows=rows, ids=ids, shape=tuple(feat.shape)) if save else None)
        mlp = getattr(self, "mlp_%d" % feat_id)
        l1, l2 = mlp[0], mlp[2]
        R = rows.shape[0]
        op1 = ops.DenseOperand(rows.view(1, R, 1, rows.shape[1]))
        h1, _ = ops.conv2d_fwd(op1, l1.pack(0), 1, 0, R, 1, bias=l1.bias, impl=ops.IMPL_SIMT)
        _, op2 = ops.norm_act_pad(h1, act=ACT_RELU, pad=0, fmt=FMT_F32)
        h2, _ = ops.conv2d_fwd(op2, l2.pack(0), 1, 0, R, 1, bias=l2.bias, impl=ops.IMPL_SIMT)
        eye = self._arange(R, feat.device)
        out, _ = ops.patch_sample_l2norm(h2, eye)
        ctx = dict(rows=rows, ids=ids, shape=tuple(feat.shape), op1=op1, h1=h1, op2=op2, h2=h2, eye=eye, feat_id=feat_id) if save else None
        return out, ctx

    def sample_bwd(self, ctx, dout):
        """dout: gradient w.r.t. the normalised rows -> gradient w.r.t. the NHWC feature map (weight grads accumulate)."""
        if not self.use_mlp:
            return ops.patch_sample_l2norm_bwd(dout, ctx["rows"], ctx["ids"], ctx["shape"])
        mlp = getattr(self, "mlp_%d" % ctx["feat_id"])
        l1, l2 = mlp[0], mlp[2]
        h1, h2 = ctx["h1"], ctx["h2"]
        R = h2.shape[1]
        dh2 = ops.patch_sample_l2norm_bwd(dout, h2.view(R, -1), ctx["eye"], tuple(h2.shape))
        d2 = ops.DenseOperand(dh2)
        ops.conv2d_wgrad(ctx["op2"], 0, d2, 0, 1, 1, R, 1, l2.weight.grad.view(l2.co, l2.ci, 1, 1), l2.bias.grad, impl=ops.IMPL_SIMT)
        da1 = ops.conv2d_dgrad_gather(dh2, l2.pack(2), 1, R, 1)
        g1, _ = ops.act_norm_bwd_reduce(tuple(h1.shape), dadd=da1, raw=h1, act=ACT_RELU)
        d1 = ops.DenseOperand(g1)
        ops.conv2d_wgrad(ctx["op1"], 0, d1, 0, 1, 1, R, 1, l1.weight.grad.view(l1.co, l1.ci, 1, 1), l1.bias.grad, impl=ops.IMPL_SIMT)
        drows = ops.conv2d_dgrad_gather(g1, l1.pack(2), 1, R, 1)
        return ops.rows_scatter_add(drows.view(R, -1), ctx["ids"], ctx["shape"])

    def _arange(self, n, device):
        tab = self.__dict__.setdefault("_eyes", {})
        key = (n, str(device))
        if key not in tab:
            tab[key] = torch.arange(n, dtype=torch.int32, device=device)
        return tab[key]

    def forward(self, feats, num_patches=64, patch_ids=None):
        return_ids, return_feats = [], []
        if self.use_mlp and not self.mlp_init:
            self.create_mlp(feats)
        if self.use_mlp:
            self.ensure_flat()
            self.refresh_packs()
        for feat_id, feat in enumerate(feats):
            _require_cuda(feat, "PatchSampleF")
            B, C_, H, W = feat.shape
            if num_patches <= 0:
                raise NotImplementedError("num_patches == 0 (dense) is outside the B200 hot path")
            if patch_ids is not None:
                patch_id = patch_ids[feat_id]
            else:
                patch_id = first_of_permutation(H * W, num_patches)
            host_ids = np.asarray(patch_id.cpu() if torch.is_tensor(patch_id) else patch_id)
            ids = torch.as_tensor(host_ids, dtype=torch.int32).to(feat.device)
            # features produced by our generators are NHWC in memory (NCHW views): no copy then
            nhwc = feat.permute(0, 2, 3, 1).contiguous().float()
            out, _ = self.sample_fwd(nhwc, ids, feat_id)
            return_ids.append(torch.as_tensor(host_ids, dtype=torch.long, device=feat.device))
            return_feats.append(out)
        return return_feats, return_ids


# ----------------------------------------------------------------------------- factories / init / schedulers
def get_scheduler(optimizer, opt):
    """networks.py:148-174."""
    if opt.lr_policy == "linear":
        def lambda_rule(epoch):
            return 1.0 - max(0, epoch + opt.epoch_count - opt.n_epochs) / float(opt.n_epochs_decay + 1)
        return lr_scheduler.LambdaLR(optimizer, lr_lambda=lambda_rule)
    if opt.lr_policy == "step":
        return lr_scheduler.StepLR(optimizer, step_size=opt.lr_decay_iters, gamma=0.1)
    if opt.lr_policy == "plateau":
        return lr_scheduler.ReduceLROnPlateau(optimizer, mode="min", factor=0.2, threshold=0.01, patience=5)
    if opt.lr_policy == "cosine":
        return lr_scheduler.CosineAnnealingLR(optimizer, T_max=opt.n_epochs, eta_min=0)
    raise NotImplementedError("learning rate policy [%s] is not implemented" % opt.lr_policy)


def init_weights(net, init_type="normal", init_gain=0.02, debug=False):
    """networks.py:191-231 (same class-name dispatch: 'Conv'/'Linear' weights, BatchNorm2d N(1, gain))."""
    def init_func(m):
        classname = m.__class__.__name__
        if hasattr(m, "weight") and (classname.find("Conv") != -1 or classname.find("Linear") != -1):
            if init_type == "normal":
                init.normal_(m.weight.data, 0.0, init_gain)
            elif init_type == "xavier":
                init.xavier_normal_(m.weight.data, gain=init_gain)
            elif init_type == "xavier_uniform":
                init.xavier_uniform_(m.weight.data, gain=1.0)
            elif init_type == "kaiming":
                init.kaiming_normal_(m.weight.data, a=0, mode="fan_in")
            elif init_type == "orthogonal":
                init.orthogonal_(m.weight.data, gain=init_gain)
            elif init_type == "none":
                m.reset_parameters()
            else:
                raise NotImplementedError("initialization method [%s] is not implemented" % init_type)
            if hasattr(m, "bias") and m.bias is not None:
                init.constant_(m.bias.data, 0.0)
        elif classname.find("BatchNorm2d") != -1:
            init.normal_(m.weight.data, 1.0, init_gain)
            init.constant_(m.bias.data, 0.0)
    net.apply(init_func)


def init_net(net, init_type="normal", init_gain=0.02, gpu_ids=[], debug=False, initialize_weights=True, gpu_idx=0):
    """networks.py:234-252."""
    if len(gpu_ids) > 0:
        assert torch.cuda.is_available()
        net.to(gpu_ids[gpu_idx])
        net.gpu_idx = gpu_idx
    if initialize_weights:
        init_weights(net, init_type, init_gain=init_gain, debug=debug)
    return net


_NORM_NAMES = ("batch", "instance", "none")


def define_G(input_nc, output_nc, ngf, netG, norm="batch", use_dropout=False, init_type="normal",
             init_gain=0.02, no_antialias=False, no_antialias_up=False, gpu_ids=[], opt=None, generate_T_imgs=False,
             num_layer_separate=0):
    """networks.py:255-325.  The B200 path builds the Resnet family with InstanceNorm and the
    antialiased resamplers (the reference's configuration for this path); other names raise the
    reference's NotImplementedError text, unsupported-but-known ones a specific one."""
    blocks = {"resnet_9blocks": 9, "resnet_6blocks": 6, "resnet_4blocks": 4}
    if netG in blocks:
        if norm != "instance":
            raise NotImplementedError("B200 path: resnet generators are built with norm='instance' (got %r)" % norm)
        if use_dropout or no_antialias or no_antialias_up:
            raise NotImplementedError("B200 path: dropout / no_antialias variants of the resnet generator are not built")
        net = ResnetGenerator(input_nc, output_nc, ngf, n_blocks=blocks[netG], opt=opt)
    elif netG == "unet256_custom":
        if norm != "instance":
            raise NotImplementedError("B200 path: unet256_custom is built with norm='instance' (got %r)" % norm)
        if use_dropout:
            raise NotImplementedError("B200 path: dropout is not built")
        net = CustomUnetGenerator(input_nc, output_nc, num_downs=8, ngf=ngf, num_layer_separate=num_layer_separate, opt=opt)
    elif netG in ("stylegan2", "smallstylegan2"):
        from .sg2_generator import StyleGAN2Generator   # forward, feature taps and explicit backward (sg2_generator.py)
        net = StyleGAN2Generator(input_nc, output_nc, ngf, use_dropout=use_dropout, n_blocks=6 if netG == "stylegan2" else 2, opt=opt)
    elif netG in ("unet_128", "unet_256"):
        raise NotImplementedError("Generator model name [%s] is on the roadmap of the B200 path but not built yet" % netG)
    else:
        raise NotImplementedError("Generator model name [%s] is not recognized" % netG)
    return init_net(net, init_type, init_gain, gpu_ids, initialize_weights=("stylegan2" not in netG))


def define_D(input_nc, ndf, netD, n_layers_D=3, norm="batch", init_type="normal", init_gain=0.02, no_antialias=False,
             num_D=3, gpu_ids=[], opt=None, gpu_idx=0):
    """networks.py:392-442."""
    if norm not in _NORM_NAMES:
        raise NotImplementedError("normalization layer [%s] is not found" % norm)
    if netD == "basic":
        net = NLayerDiscriminator(input_nc, ndf, n_layers=3, norm=norm)
    elif netD == "n_layers":
        net = NLayerDiscriminator(input_nc, ndf, n_layers_D, norm=norm)
    elif netD == "multiscale":
        net = MultiscaleDiscriminator(input_nc, ndf, n_layers=n_layers_D, norm=norm, num_D=num_D, opt=opt)
    elif netD == "pixel" or "stylegan2" in netD:
        raise NotImplementedError("Discriminator model name [%s] is not built on the B200 path" % netD)
    else:
        raise NotImplementedError("Discriminator model name [%s] is not recognized" % netD)
    return init_net(net, init_type, init_gain, gpu_ids, initialize_weights=("stylegan2" not in netD), gpu_idx=gpu_idx)


def define_F(input_nc, netF, norm="batch", use_dropout=False, init_type="normal", init_gain=0.02, no_antialias=False,
             gpu_ids=[], opt=None):
    """networks.py:328-341."""
    if netF == "sample":
        net = PatchSampleF(use_mlp=False, init_type=init_type, init_gain=init_gain, gpu_ids=gpu_ids, nc=opt.netF_nc)
    elif netF == "mlp_sample":
        net = PatchSampleF(use_mlp=True, init_type=init_type, init_gain=init_gain, gpu_ids=gpu_ids, nc=opt.netF_nc)
    elif netF in ("global_pool", "reshape", "strided_conv"):
        raise NotImplementedError("projection model name [%s] is not built on the B200 path" % netF)
    else:
        raise NotImplementedError("projection model name [%s] is not recognized" % netF)
    return init_net(net, init_type, init_gain, gpu_ids)
