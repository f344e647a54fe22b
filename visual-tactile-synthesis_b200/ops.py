"""Tensor-level wrappers over the C ABI (one function per entry point of include/skit_b200.h).

Device layouts: feature maps NHWC fp32; conv operands NHWC with halo, fp32 or
bf16 hi/lo planes; images/patches NCHW fp32.  PyTorch only provides memory and streams here.
"""
import ctypes as C

import torch

from . import _lib as L
from ._lib import (ACT_LRELU, ACT_NONE, ACT_RELU, FMT_BF16X2, FMT_F32, IMPL_AUTO, IMPL_SIMT, IMPL_TC,
                   NORM_BATCH, NORM_INSTANCE, NORM_NONE, PAD_REFLECT, PAD_REPLICATE, PAD_ZERO)

_p = L.ptr


class ZeroArena:
    """Zero-initialised small scratch of one train step (norm statistics of every conv, the sums of every norm backward):
    ~170 tiny memsets on the critical path become one.  (The split-K partials of the weight gradients are large and
    live on the side stream: they keep their own memsets.)  Slices are bump-allocated in launch order — deterministic from
    step to step — and `begin()` clears exactly the bytes the previous step used; a slice beyond that mark is cleared
    on its own.  Only active inside SinSKITGModel._step_body; everywhere else `zeros()` is torch.zeros."""

    def __init__(self, device, nbytes=8 << 20):
        self.buf = torch.empty(nbytes, dtype=torch.uint8, device=device)
        self.off, self.clean, self.high = 0, 0, 0

    def begin(self):
        if self.high:
            zero_(self.buf[:self.high])
        self.clean, self.off = self.high, 0

    def take(self, shape, dtype):
        n = 1
        for d in shape:
            n *= int(d)
        nbytes = n * torch.empty((), dtype=dtype).element_size()
        start, end = self.off, self.off + ((nbytes + 255) // 256) * 256
        if end > self.buf.numel():
            return None
        view = self.buf[start:start + nbytes].view(dtype).view(shape)
        if end > self.clean:
            zero_(view)
        self.off = end
        self.high = max(self.high, end)
        return view


SUM_REPLICAS = 8   # SKIT_SUM_REPLICAS of include/skit_b200.h (tests/test_abi.py checks the two agree)
ARENA = None   # the active ZeroArena (set by the model around a train step)


def zero_(t):
    """t.zero_() as a memset (skit_zero_bytes): no fill kernel, a memset node inside a captured graph."""
    L.call("skit_zero_bytes", _p(t), t.numel() * t.element_size(), L.stream())
    L.launches -= 1     # a memset, not a kernel
    return t


def zeros_big(shape, dtype, device):
    """torch.zeros for the large scratch / scatter buffers of the backward pass, through a memset."""
    return zero_(torch.empty(tuple(shape), dtype=dtype, device=device))


def zeros(shape, dtype, device):
    if ARENA is not None:
        t = ARENA.take(tuple(shape), dtype)
        if t is not None:
            return t
    return zeros_big(shape, dtype, device)


class Operand:
    """A haloed NHWC conv operand on the device (fp32, or bf16 hi/lo planes for tcgen05)."""

    def __init__(self, n, h, w, c, pad, fmt, device):
        self.n, self.h, self.w, self.c, self.pad, self.fmt = n, h, w, c, pad, fmt
        self.hp, self.wp = h + 2 * pad, w + 2 * pad
        if fmt == FMT_F32:
            self.data = torch.empty((n, self.hp, self.wp, c), dtype=torch.float32, device=device)
            p0, p1 = self.data.data_ptr(), None
        else:
            self.data = torch.empty((2, n, self.hp, self.wp, c), dtype=torch.bfloat16, device=device)
            p0, p1 = self.data[0].data_ptr(), self.data[1].data_ptr()
        self.struct = L.SkitOperand(p0, p1, fmt, n, self.hp, self.wp, c)

    def ref(self):
        return C.byref(self.struct)

    def to_float(self):
        """fp32 view of the contents (tests only)."""
        return self.data if self.fmt == FMT_F32 else self.data[0].float() + self.data[1].float()


class PackedWeights:
    """Per-step repack of one conv's reference-layout weight (skit_pack_conv_weights)."""

    def __init__(self, w, mode, want_f32=True, want_bf16=False, kpad=0, cp=0):
        co, ci, k, _ = w.shape
        self.k, self.mode, self.cp = k, mode, cp
        # the GEMM reduces over `rci` channels per tap and produces `rco` columns
        self.rci, self.rco = (ci, co) if mode in (0, 4) else (co, ci)
        self.kpad, self.kw = 0, 0
        if mode in (4, 5):               # x-folded bf16 pack [k][N][64]: the filter columns live in the channel axis
            assert want_bf16 and not want_f32 and cp >= self.rci and k * cp <= 64
            self.rci, self.kw, self.kpad = 64, 1, 64
        elif kpad and kpad > self.rci:     # bf16 pack with a zero-padded reduction axis (thin layers on tcgen05)
            assert want_bf16 and not want_f32 and kpad % 8 == 0
            self.kpad, self.rci = kpad, kpad
        dev = w.device
        taps = k if mode in (4, 5) else k * k
        self.f32 = torch.empty((taps * self.rci, self.rco), dtype=torch.float32, device=dev) if want_f32 else None
        self.hi = torch.empty((taps, self.rco, self.rci), dtype=torch.bfloat16, device=dev) if want_bf16 else None
        self.lo = torch.empty_like(self.hi) if want_bf16 else None
        self.co, self.ci = co, ci
        self.struct = L.SkitWeights(self.f32.data_ptr() if want_f32 else None,
                                    self.hi.data_ptr() if want_bf16 else None,
                                    self.lo.data_ptr() if want_bf16 else None, k, self.rci, self.rco, self.kw)
        self.refresh(w)

    def refresh(self, w):
        if self.mode in (4, 5):
            L.call("skit_pack_conv_weights_folded", _p(w.detach()), self.co, self.ci, self.k, self.mode, self.cp,
                   _p(self.hi), _p(self.lo), L.stream())
            return
        if self.kpad:
            L.call("skit_pack_conv_weights_padded", _p(w.detach()), self.co, self.ci, self.k, self.mode, self.kpad,
                   _p(self.hi), _p(self.lo), L.stream())
            return
        L.call("skit_pack_conv_weights", _p(w.detach()), self.co, self.ci, self.k, self.mode,
               _p(self.f32), _p(self.hi), _p(self.lo), L.stream())

    def ref(self):
        return C.byref(self.struct)

    def desc(self, w, start):
        """Descriptor of this pack for the batched refresh (skit_pack_conv_weights_batched)."""
        return L.SkitPackDesc(w.data_ptr(), self.f32.data_ptr() if self.f32 is not None else None,
                              self.hi.data_ptr() if self.hi is not None else None,
                              self.lo.data_ptr() if self.lo is not None else None, start,
                              self.co, self.ci, self.k, self.mode, self.kpad, self.cp)

    def numel(self):
        return (self.k if self.mode in (4, 5) else self.k * self.k) * self.rco * self.rci

    def tiles(self):
        """32x32 (N x K) tiles of the shared-memory staged refresh; 0 when the pack is not eligible for it."""
        if self.mode in (4, 5) or self.k > 4 or (self.hi is not None and self.rci % 2):
            return 0
        return -(-self.rco // 32) * -(-self.rci // 32)


class PackTable:
    """All packs of one net refreshed by at most two kernel launches: the k <= 4 packs through the shared-memory tiled
    kernel (skit_pack_conv_weights_tiled), the rest (7x7, x-folded) element-wise (skit_pack_conv_weights_batched)."""

    def __init__(self, pairs, device):
        """pairs: [(weight tensor [co,ci,k,k], PackedWeights)]"""
        tiled = [(w, pk) for w, pk in pairs if pk.tiles() > 0]
        flat = [(w, pk) for w, pk in pairs if pk.tiles() == 0]
        self.tiled = self._table(tiled, lambda pk: pk.tiles(), device)
        self.flat = self._table(flat, lambda pk: pk.numel(), device)
        self.max_k = max([pk.k for _, pk in tiled], default=0)
        self.key = tuple((w.data_ptr(), id(pk)) for w, pk in pairs)

    @staticmethod
    def _table(pairs, size, device):
        if not pairs:
            return None
        arr = (L.SkitPackDesc * len(pairs))()
        start = 0
        for i, (w, pk) in enumerate(pairs):
            arr[i] = pk.desc(w, start)
            start += size(pk)
        raw = torch.frombuffer(bytearray(bytes(arr)), dtype=torch.uint8).clone()
        return raw.to(device), len(pairs), start

    def refresh(self):
        if self.tiled is not None:
            dev, n, total = self.tiled
            L.call("skit_pack_conv_weights_tiled", _p(dev), n, total, self.max_k, L.stream())
        if self.flat is not None:
            dev, n, total = self.flat
            L.call("skit_pack_conv_weights_batched", _p(dev), n, total, L.stream())


def conv2d_fwd(x, w, stride, org, ho, wo, bias=None, stats_mode=NORM_NONE, impl=IMPL_AUTO, out=None):
    """Valid conv over a haloed operand -> (y NHWC fp32, stats double [groups, co, 2] or None)."""
    co = w.rco
    dev = x.data.device
    y = out if out is not None else torch.empty((x.n, ho, wo, co), dtype=torch.float32, device=dev)
    stats = None
    if stats_mode != NORM_NONE:
        groups = x.n if stats_mode == NORM_INSTANCE else 1
        stats = zeros((groups, co, 2), torch.float64, dev)
    L.call("skit_conv2d_fwd", x.ref(), w.ref(), stride, org, ho, wo, _p(bias), _p(y), _p(stats), stats_mode, impl, L.stream())
    return y, stats


def conv2d_dgrad_gather(dy, wg, stride, hp, wp):
    n, ho, wo, co = dy.shape
    dx = torch.empty((n, hp, wp, wg.rco), dtype=torch.float32, device=dy.device)
    L.call("skit_conv2d_dgrad_gather", _p(dy), n, ho, wo, co, wg.ref(), stride, hp, wp, _p(dx), L.stream())
    return dx


def set_backward_terms(terms):
    """2 (default): backward tensor-core launches drop one hi/lo cross term; 3: the forward's full three-term product."""
    L.call("skit_set_backward_terms", int(terms))
    L.launches -= 1     # not a kernel launch
    backward_terms.current = int(terms)


class backward_terms:
    """`with backward_terms(3): net.bwd(...)` — the launches issued inside use that backward precision (host-side switch read
    at launch time, so it is safe with side streams and graph capture), the previous setting is restored on exit."""
    current = 3 if __import__("os").environ.get("SKIT_BWD_TERMS", "2").startswith("3") else 2   # the library's own default

    def __init__(self, terms):
        self.terms = terms

    def __enter__(self):
        self.prev = backward_terms.current
        if self.terms != self.prev:
            set_backward_terms(self.terms)

    def __exit__(self, *a):
        if self.terms != self.prev:
            set_backward_terms(self.prev)


def conv2d_dgrad_s1(dy_op, w1):
    """Stride-1 input gradient on tcgen05: dy_op carries a zero halo of k-1; w1 = mode-1 pack -> dx NHWC fp32."""
    k = w1.k
    dx = torch.empty((dy_op.n, dy_op.hp - k + 1, dy_op.wp - k + 1, w1.rco), dtype=torch.float32, device=dy_op.data.device)
    L.call("skit_conv2d_dgrad_s1", dy_op.ref(), w1.ref(), _p(dx), L.stream())
    return dx


def conv2d_dgrad_s2(dy_op, dy_pad, wp, k, ho, wo, hp, wp_):
    """Stride-2 input gradient on tcgen05 (four parity sub-convolutions); dy_op: bf16x2, halo k/2-1."""
    dx = torch.empty((dy_op.n, hp, wp_, wp.rco), dtype=torch.float32, device=dy_op.data.device)
    L.call("skit_conv2d_dgrad_s2", dy_op.ref(), dy_pad, wp.ref(), k, ho, wo, hp, wp_, _p(dx), L.stream())
    return dx


def conv2d_wgrad(x, org, dy, dy_org, k, stride, ho, wo, dw, dbias=None, impl=IMPL_AUTO):
    """Accumulates into dw ([co][ci][k][k] view of the flat grad bucket) and dbias.  Operands may carry zero padding
    channels beyond dw's real (co, ci)."""
    co, ci = dy.c, x.c
    scratch = zeros_big((k * k * ci * co,), torch.float32, x.data.device)   # big, and off the critical path
    L.call("skit_conv2d_wgrad_ex", x.ref(), org, dy.ref(), dy_org, k, stride, ho, wo, _p(scratch), _p(dw), _p(dbias), impl,
           int(dw.shape[0]), int(dw.shape[1]), L.stream())


def fold_x(thin, kw):
    """thin: haloed Operand with few channels -> bf16x2 Operand [n, hp, wp-kw+1, 64] whose channel axis holds the kw shifted
    copies (skit_fold_x_operand): a (k x kw) conv over `thin` = a (k x 1) conv over the result with a folded pack."""
    f = Operand(thin.n, thin.hp, thin.wp - kw + 1, 64, 0, FMT_BF16X2, thin.data.device)
    L.call("skit_fold_x_operand", thin.ref(), kw, f.ref(), L.stream())
    return f


def conv2d_wgrad_folded(xf, dy, dy_org, k, kw, cp, ho, wo, dw):
    """Weight gradient against an x-folded input operand; accumulates into dw [co][ci][k][kw]."""
    scratch = zeros_big((k * 64 * dy.c,), torch.float32, xf.data.device)
    L.call("skit_conv2d_wgrad_folded", xf.ref(), 0, dy.ref(), dy_org, k, kw, cp, ho, wo, _p(scratch), _p(dw),
           int(dw.shape[0]), int(dw.shape[1]), L.stream())


def conv2d_wgrad_dyfolded(x, dyf, k, cp, ho, wo, dw):
    """Weight gradient of a thin-output k x k layer against its x-folded gradient operand; accumulates into dw [co][ci][k][k]."""
    scratch = zeros_big((k * 64 * x.c,), torch.float32, x.data.device)
    L.call("skit_conv2d_wgrad_dyfolded", x.ref(), dyf.ref(), k, cp, ho, wo, _p(scratch), _p(dw), int(dw.shape[0]), int(dw.shape[1]), L.stream())


def dbias_n(dy_op, org, ho, wo, nch, db):
    L.call("skit_dbias_n", dy_op.ref(), org, ho, wo, nch, _p(db), L.stream())


def stats_finalize(stats, count, eps=1e-5, running_mean=None, running_var=None, momentum=0.1):
    groups, c, _ = stats.shape
    mr = torch.empty((groups, c, 2), dtype=torch.float32, device=stats.device)
    L.call("skit_stats_finalize", _p(stats), groups, c, float(count), eps, _p(mr), _p(running_mean), _p(running_var), momentum, L.stream())
    return mr


def norm_act_pad(raw, mr=None, norm_mode=NORM_NONE, gamma=None, beta=None, act=ACT_NONE, residual=None,
                 want_dense=False, pad=0, pad_mode=PAD_ZERO, fmt=None):
    """-> (dense NHWC fp32 or None, Operand or None)."""
    n, h, w, c = raw.shape
    dense = torch.empty_like(raw) if want_dense else None
    op = Operand(n, h, w, c, pad, fmt, raw.device) if fmt is not None else None
    L.call("skit_norm_act_pad", _p(raw), n, h, w, c, _p(mr), norm_mode, _p(gamma), _p(beta), act, _p(residual), _p(dense),
           op.ref() if op is not None else None, pad, pad_mode, L.stream())
    return dense, op


def norm_act_pad_stats(raw, stats, count, norm_mode, gamma=None, beta=None, act=ACT_NONE, residual=None,
                       want_dense=False, pad=0, pad_mode=PAD_ZERO, fmt=None, eps=1e-5):
    """norm_act_pad fed by the conv epilogue's fp64 (sum, sum of squares) directly: the statistics are finalised inside the
    pass itself -> (dense or None, Operand or None, mean_rstd [groups, c, 2] for the backward)."""
    n, h, w, c = raw.shape
    dense = torch.empty_like(raw) if want_dense else None
    op = Operand(n, h, w, c, pad, fmt, raw.device) if fmt is not None else None
    mr = torch.empty((stats.shape[0], c, 2), dtype=torch.float32, device=raw.device)
    L.call("skit_norm_act_pad_stats", _p(raw), n, h, w, c, _p(stats), float(count), eps, _p(mr), norm_mode, _p(gamma), _p(beta), act,
           _p(residual), _p(dense), op.ref() if op is not None else None, 0, pad, pad_mode, L.stream())
    return dense, op, mr


def act_norm_bwd_reduce(shape, dpad=None, pad=0, pad_mode=PAD_ZERO, dadd=None, raw=None, mr=None, norm_mode=NORM_NONE,
                        gamma=None, beta=None, act=ACT_NONE):
    """-> (g NHWC fp32, sums double [groups, c, 2] or None)."""
    n, h, w, c = shape
    dev = (dpad if dpad is not None else dadd).device
    g = torch.empty((n, h, w, c), dtype=torch.float32, device=dev)
    sums = None
    if norm_mode != NORM_NONE:
        sums = zeros(((2 * SUM_REPLICAS + 1) * (n if norm_mode == NORM_INSTANCE else 1) * c + 1,), torch.float64, dev)
    L.call("skit_act_norm_bwd_reduce", _p(dpad), pad, pad_mode, _p(dadd), _p(raw), n, h, w, c, _p(mr), norm_mode,
           _p(gamma), _p(beta), act, _p(g), _p(sums), L.stream())
    return g, sums


def norm_act_pad_into(op, c_off, raw, mr=None, norm_mode=NORM_NONE, act=ACT_NONE, pad=0, pad_mode=PAD_ZERO):
    """Fill channels [c_off, c_off + c) of an existing haloed operand (decoder concat, slice by slice)."""
    n, h, w, c = raw.shape
    L.call("skit_norm_act_pad_ex", _p(raw), n, h, w, c, _p(mr), norm_mode, None, None, act, None, None,
           op.ref(), c_off, pad, pad_mode, L.stream())


def act_norm_bwd_reduce_ex(shape, dpad=None, pad=0, pad_mode=PAD_ZERO, dadd=None, dadd2=None, dadd_c0=0, dadd_ctot=None,
                           dadd_relu_mask=False, raw=None, mr=None, norm_mode=NORM_NONE, act=ACT_NONE, gamma=None, beta=None,
                           want_dsum=False):
    """act_norm_bwd_reduce with sliced / doubled / ReLU-masked dense gradients -> (g, sums) or (g, sums, dsum) where
    dsum = fold(dpad) + dadd (+ dadd2) is the incoming gradient itself (before act')."""
    n, h, w, c = shape
    dev = (dpad if dpad is not None else dadd).device
    g = torch.empty((n, h, w, c), dtype=torch.float32, device=dev)
    # without an activation the masked gradient g IS the incoming gradient: one tensor serves both (a ResnetBlock's second conv)
    alias = want_dsum and act == ACT_NONE and not dadd_relu_mask
    dsum = torch.empty((n, h, w, c), dtype=torch.float32, device=dev) if (want_dsum and not alias) else None
    sums = None
    if norm_mode != NORM_NONE:
        sums = zeros(((2 * SUM_REPLICAS + 1) * (n if norm_mode == NORM_INSTANCE else 1) * c + 1,), torch.float64, dev)
    L.call("skit_act_norm_bwd_reduce_ex2", _p(dpad), pad, pad_mode, _p(dadd), _p(dadd2), dadd_c0,
           c if dadd_ctot is None else dadd_ctot, int(dadd_relu_mask), _p(raw), n, h, w, c, _p(mr), norm_mode,
           _p(gamma), _p(beta), act, _p(g), _p(sums), _p(dsum), L.stream())
    return (g, sums, g if alias else dsum) if want_dsum else (g, sums)


def conv_transpose2d_fwd(x, wg, stride, pad, bias=None, stats_mode=NORM_NONE, out=None, out_c0=0):
    """x: dense NHWC fp32; wg: mode-2 PackedWeights of the [ci][co][k][k] weight -> (y NHWC (or `out` slice), stats)."""
    n, h, w, ci = x.shape
    co, k = wg.rco, wg.k
    ho, wo = (h - 1) * stride - 2 * pad + k, (w - 1) * stride - 2 * pad + k
    y = out if out is not None else torch.empty((n, ho, wo, co), dtype=torch.float32, device=x.device)
    stats = None
    if stats_mode != NORM_NONE:
        stats = zeros((n if stats_mode == NORM_INSTANCE else 1, co, 2), torch.float64, x.device)
    L.call("skit_conv_transpose2d_fwd", _p(x), n, h, w, ci, wg.ref(), stride, pad, ho, wo, _p(bias), _p(y), y.shape[3], out_c0,
           _p(stats), stats_mode, L.stream())
    return y, stats


def dbias(dy_op, org, ho, wo, db):
    L.call("skit_dbias", dy_op.ref(), org, ho, wo, _p(db), L.stream())


def g_head_bwd_split(raw, mask, dI, dT, pad):
    n, h, w, _ = raw.shape
    opI = Operand(n, h, w, 3, pad, FMT_F32, raw.device)
    opT = Operand(n, h, w, 2, pad, FMT_F32, raw.device)
    L.call("skit_g_head_bwd_split", _p(raw), _p(mask), _p(dI), _p(dT), n, h, w, opI.ref(), opT.ref(), pad, L.stream())
    return opI, opT


def norm_bwd_apply(g, raw=None, mr=None, norm_mode=NORM_NONE, gamma=None, sums=None, count=0.0, dgamma=None, dbeta=None,
                   pad=0, fmt=FMT_F32, extra=None):
    n, h, w, c = g.shape
    op = Operand(n, h, w, c, pad, fmt, g.device)
    L.call("skit_norm_bwd_apply_ex", _p(g), _p(raw), n, h, w, c, _p(mr), norm_mode, _p(gamma), _p(sums), float(count),
           _p(dgamma), _p(dbeta), _p(extra), op.ref(), pad, L.stream())
    return op


def mask_mul_(x, m):
    """x[n,c,h,w] *= m[n,1,h,w] in place."""
    n, c, h, w = x.shape
    L.call("skit_mask_mul", _p(x), _p(m), n, c, h, w, L.stream())
    return x


def channel_mean(x):
    n, c, h, w = x.shape
    y = torch.empty((n, 1, h, w), dtype=torch.float32, device=x.device)
    L.call("skit_channel_mean", _p(x), n, c, h, w, _p(y), L.stream())
    return y


def channel_mean_bwd(dy, dx):
    n, c, h, w = dx.shape
    L.call("skit_channel_mean_bwd", _p(dy), n, c, h, w, _p(dx), L.stream())


def _resample(name, x, out_hw):
    n, h, w, c = x.shape
    y = torch.empty((n, out_hw[0], out_hw[1], c), dtype=torch.float32, device=x.device)
    return n, h, w, c, y


def blur_down_fwd(x):
    n, h, w, c = x.shape
    y = torch.empty((n, h // 2, w // 2, c), dtype=torch.float32, device=x.device)
    L.call("skit_blur_down_fwd", _p(x), n, h, w, c, _p(y), L.stream())
    return y


def blur_down_bwd(dy, h, w):
    n, _, _, c = dy.shape
    dx = torch.empty((n, h, w, c), dtype=torch.float32, device=dy.device)
    L.call("skit_blur_down_bwd", _p(dy), n, h, w, c, _p(dx), L.stream())
    return dx


def blur_up_fwd(x):
    n, h, w, c = x.shape
    y = torch.empty((n, 2 * h, 2 * w, c), dtype=torch.float32, device=x.device)
    L.call("skit_blur_up_fwd", _p(x), n, h, w, c, _p(y), L.stream())
    return y


def blur_up_bwd(dy):
    n, h2, w2, c = dy.shape
    dx = torch.empty((n, h2 // 2, w2 // 2, c), dtype=torch.float32, device=dy.device)
    L.call("skit_blur_up_bwd", _p(dy), n, h2 // 2, w2 // 2, c, _p(dx), L.stream())
    return dx


def _src_arrays(srcs):
    k = len(srcs)
    ptrs = (C.c_void_p * k)(*[s.data_ptr() for s in srcs])
    chans = (C.c_int * k)(*[int(s.shape[1]) for s in srcs])
    return ptrs, chans


def nchw_cat_to_operand(srcs, pad, pad_mode, fmt=FMT_F32, cpad=0):
    """srcs: list of contiguous NCHW fp32 tensors with equal N,H,W -> Operand of the channel concat (fp32 or bf16x2),
    zero-padded to `cpad` channels when given."""
    for s in srcs:
        assert s.is_cuda and s.is_contiguous() and s.dtype == torch.float32
    n, _, h, w = srcs[0].shape
    ctot = sum(int(s.shape[1]) for s in srcs)
    op = Operand(n, h, w, max(ctot, cpad), pad, fmt, srcs[0].device)
    ptrs, chans = _src_arrays(srcs)
    L.call("skit_nchw_cat_to_operand", ptrs, chans, len(srcs), n, h, w, op.ref(), pad, pad_mode, L.stream())
    return op


def operand_grad_to_nchw(dpad, h, w, pad, pad_mode, c0, cs, dst=None, accumulate=False):
    n, _, _, c = dpad.shape
    if dst is None:
        dst = torch.empty((n, cs, h, w), dtype=torch.float32, device=dpad.device)
        accumulate = False
    L.call("skit_operand_grad_to_nchw", _p(dpad), n, h, w, c, pad, pad_mode, c0, cs, _p(dst), int(accumulate), L.stream())
    return dst


def g_head_fwd(raw, mask, scale_nz=0.25, want_normal=True):
    n, h, w, _ = raw.shape
    dev = raw.device
    fI = torch.empty((n, 3, h, w), dtype=torch.float32, device=dev)
    fT = torch.empty((n, 2, h, w), dtype=torch.float32, device=dev)
    fN = torch.empty((n, 3, h, w), dtype=torch.float32, device=dev) if want_normal else None
    L.call("skit_g_head_fwd", _p(raw), _p(mask), n, h, w, scale_nz, _p(fI), _p(fT), _p(fN), L.stream())
    return fI, fT, fN


def g_head_bwd(raw, mask, dI, dT, pad, fmt=FMT_F32, cpad=5):
    n, h, w, _ = raw.shape
    op = Operand(n, h, w, max(5, cpad), pad, fmt, raw.device)
    L.call("skit_g_head_bwd", _p(raw), _p(mask), _p(dI), _p(dT), n, h, w, op.ref(), pad, L.stream())
    return op


def diffaug_bs_mask(x, mask, u_b, u_s):
    n, _, h, w = x.shape
    y = torch.empty_like(x)
    L.call("skit_diffaug_bs_mask", _p(x), _p(mask), _p(u_b), _p(u_s), n, h, w, _p(y), L.stream())
    return y


def avgpool3s2_fwd(x):
    n, c, h, w = x.shape
    y = torch.empty((n, c, (h - 1) // 2 + 1, (w - 1) // 2 + 1), dtype=torch.float32, device=x.device)
    L.call("skit_avgpool3s2_fwd", _p(x), n * c, h, w, _p(y), L.stream())
    return y


def avgpool3s2_bwd(dy, h, w, dx=None, accumulate=False):
    n, c = dy.shape[:2]
    if dx is None:
        dx = torch.empty((n, c, h, w), dtype=torch.float32, device=dy.device)
        accumulate = False
    L.call("skit_avgpool3s2_bwd", _p(dy), n * c, h, w, _p(dx), int(accumulate), L.stream())
    return dx


def patch_gather(srcs, ox, oy, ps, ctot=None, coffs=None, dst=None):
    """srcs: [1,C,H,W] NCHW tensors; ox/oy int32 device tensors [P] -> [P, ctot, ps, ps]."""
    for s in srcs:
        assert s.shape[0] == 1, "coords should have batch size of 1"  # reference: model_utils.py:235
    h, w = srcs[0].shape[-2:]
    npatch = int(ox.numel())
    total = sum(int(s.shape[1]) for s in srcs)
    ctot = total if ctot is None else ctot
    if dst is None:
        dst = torch.empty((npatch, ctot, ps, ps), dtype=torch.float32, device=srcs[0].device)
    ptrs, chans = _src_arrays(srcs)
    co = None if coffs is None else (C.c_int * len(srcs))(*coffs)
    L.call("skit_patch_gather", ptrs, chans, co, len(srcs), h, w, _p(ox), _p(oy), npatch, ps, _p(dst), ctot, L.stream())
    return dst


def patch_scatter_add(dpatch, coff, cs, ox, oy, dsrc):
    npatch, ctot, ps, _ = dpatch.shape
    h, w = dsrc.shape[-2:]
    L.call("skit_patch_scatter_add", _p(dpatch), ctot, coff, cs, h, w, _p(ox), _p(oy), npatch, ps, _p(dsrc), L.stream())


def gan_softplus(pred, sign, loss, dpred=None, gscale=0.0):
    """pred: [N, h, w, 1] (or any [N, ...]); loss: [N] accumulator."""
    n = pred.shape[0]
    hw = pred.numel() // n
    L.call("skit_gan_softplus", _p(pred), n, hw, float(sign), _p(loss), _p(dpred), float(gscale), L.stream())


GAN_MODES = {"nonsaturating": 0, "hinge": 1, "wgan": 2, "wgangp": 2, "lsgan": 3, "vanilla": 4}


def gan_loss(pred, mode, target_is_real, target, loss, dpred=None, gscale=0.0):
    """GANLoss for one scale in any of the reference's modes (networks.py:500-522); loss: [N] per-sample accumulator."""
    n = pred.shape[0]
    hw = pred.numel() // n
    L.call("skit_gan_loss", _p(pred), n, hw, GAN_MODES[mode], int(bool(target_is_real)), float(target), _p(loss), _p(dpred), float(gscale),
           L.stream())


def l1_loss(a, b, scale, loss, grad=None, gscale=0.0, accumulate=False):
    L.call("skit_l1_loss", _p(a), _p(b), a.numel(), float(scale), _p(loss), _p(grad), float(gscale), int(accumulate), L.stream())


def adam_step(p, g, m, v, step, lr, beta1, beta2, eps=1e-8, grad_scale=1.0):
    L.call("skit_adam_step", _p(p), _p(g), _p(m), _p(v), p.numel(), int(step), float(lr), float(beta1), float(beta2),
           float(eps), float(grad_scale), L.stream())


def adam_step_dev(p, g, m, v, hyper, beta1, beta2, eps=1e-8, grad_scale=1.0):
    """hyper: device fp32 [2] = [lr / bias_correction1, 1 / sqrt(bias_correction2)] (graph-replay safe)."""
    L.call("skit_adam_step_dev", _p(p), _p(g), _p(m), _p(v), p.numel(), _p(hyper), float(beta1), float(beta2),
           float(eps), float(grad_scale), L.stream())


def patch_sample_l2norm(feat_nhwc, ids, keep_pre=False):
    b, h, w, c = feat_nhwc.shape
    npatch = int(ids.numel())
    out = torch.empty((b * npatch, c), dtype=torch.float32, device=feat_nhwc.device)
    pre = torch.empty_like(out) if keep_pre else None
    L.call("skit_patch_sample_l2norm", _p(feat_nhwc), b, h * w, c, _p(ids), npatch, _p(out), _p(pre), L.stream())
    return out, pre


def patch_sample_l2norm_bwd(dout, pre, ids, feat_shape):
    b, h, w, c = feat_shape
    dfeat = zeros_big(feat_shape, torch.float32, dout.device)
    L.call("skit_patch_sample_l2norm_bwd", _p(dout), _p(pre), b, h * w, c, _p(ids), int(ids.numel()), _p(dfeat), L.stream())
    return dfeat


def rows_scatter_add(drows, ids, feat_shape):
    b, h, w, c = feat_shape
    dfeat = zeros_big(feat_shape, torch.float32, drows.device)
    L.call("skit_rows_scatter_add", _p(drows), b, h * w, c, _p(ids), int(ids.numel()), _p(dfeat), L.stream())
    return dfeat


class DenseOperand:
    """An existing dense NHWC fp32 tensor viewed as a halo-free conv operand (no copy)."""

    def __init__(self, t):
        n, h, w, c = t.shape
        assert t.is_contiguous() and t.dtype == torch.float32
        self.data, self.n, self.h, self.w, self.c, self.pad, self.fmt = t, n, h, w, c, 0, FMT_F32
        self.hp, self.wp = h, w
        self.struct = L.SkitOperand(t.data_ptr(), None, FMT_F32, n, h, w, c)

    def ref(self):
        return C.byref(self.struct)


def patchnce(q, k, b, nce_T, want_grad=False, gscale=1.0):
    rows, dim = q.shape
    npatch = rows // b
    loss = torch.empty((rows,), dtype=torch.float32, device=q.device)
    dq = torch.empty_like(q) if want_grad else None
    L.call("skit_patchnce", _p(q), _p(k), b, npatch, dim, 1.0 / nce_T, _p(loss), _p(dq), float(gscale), L.stream())
    return loss, dq
