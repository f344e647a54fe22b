"""Patch / normal / positional-encoding utilities of the hot path, mirroring
models/model_utils.py (find_coords_for_patch :23-69, get_patch_in_input :72-405, compute_normal
:408-428) and thirdparty/mmgeneration/positional_encoding.py (make_grid2d :113-159).
Coordinate arithmetic stays on the host in float64 NumPy exactly like the reference; the gather
itself is one coalesced CUDA kernel over all patches (no `input.repeat(NT,1,1,1)`)."""
import math
import random as _pyrandom

import numpy as np
import torch

from . import ops


def find_coords_for_patch(coords, scale_multiplier=1):
    """model_utils.py:37-57: rows [ROI_x, ROI_y, ROI_h, ROI_w, patch_crop_size, resize_ratio, crop_x, crop_y];
    offset = round((ROI + crop / ratio) * mult), cutout = round(patch_crop_size / ratio * mult);
    np.round (half to even) in float64, then float32 -> int32."""
    c = np.squeeze(np.asarray(coords, dtype=np.float64))
    if c.ndim == 1:
        c = c[None]
    ox = np.round((c[..., 0] + c[..., -2] / c[..., -3]) * scale_multiplier)
    oy = np.round((c[..., 1] + c[..., -1] / c[..., -3]) * scale_multiplier)
    cs = np.round(c[..., -4] / c[..., -3] * scale_multiplier)
    f = lambda a: np.asarray(a, dtype=np.float32).astype(np.int32)
    return f(ox), f(oy), f(cs)


class _OffsetTable:
    """Candidate (row, col) offsets of the random-patch mode (model_utils.py:212-222): the nonzero positions, in
    row-major order, of the dilated mask `box` ([oh, ow] bool; None = every position).  Only the sampled entries are ever
    materialised: the k-th candidate is found through the per-row cumulative counts."""

    def __init__(self, box, oh, ow):
        self.box, self.oh, self.ow = box, oh, ow
        if box is None:
            self.n = oh * ow
        else:
            self.rowcum = np.cumsum(box.sum(axis=1, dtype=np.int64))
            self.n = int(self.rowcum[-1]) if oh > 0 else 0

    def __len__(self):
        return self.n

    def at(self, picks):
        picks = np.asarray(picks, dtype=np.int64)
        if self.box is None:
            return np.divmod(picks, self.ow)
        rows = np.searchsorted(self.rowcum, picks, side="right")
        before = np.where(rows > 0, self.rowcum[np.maximum(rows - 1, 0)], 0)
        cols = np.array([np.flatnonzero(self.box[r])[k] for r, k in zip(rows, picks - before)], dtype=np.int64)
        return rows, cols

    @property
    def rows(self):
        return self.at(np.arange(self.n))[0] if self.box is None else np.nonzero(self.box)[0]

    @property
    def cols(self):
        return self.at(np.arange(self.n))[1] if self.box is None else np.nonzero(self.box)[1]

    def sample(self, k, rng=_pyrandom):
        pick = rng.sample(range(len(self)), k)          # the reference's `random.sample(range(nnz), NF)` draw
        rows, cols = self.at(pick)
        return cols.astype(np.int32), rows.astype(np.int32)  # (offset_x, offset_y)


class _BitOffsetTable(_OffsetTable):
    """The same table from the device kernel's output: one bit per position (row-major, little-endian within each 32-bit word) and
    the per-row counts; only the sampled rows are ever unpacked."""

    def __init__(self, bits, rowcount, oh, ow):
        self.bits, self.oh, self.ow = bits, oh, ow           # bits: uint8 [oh, 4 * words]
        self.box = None
        self.rowcum = np.cumsum(rowcount.astype(np.int64))
        self.n = int(self.rowcum[-1]) if oh > 0 else 0

    def _row(self, r):
        return np.flatnonzero(np.unpackbits(self.bits[r], bitorder="little")[:self.ow])

    def at(self, picks):
        picks = np.asarray(picks, dtype=np.int64)
        rows = np.searchsorted(self.rowcum, picks, side="right")
        before = np.where(rows > 0, self.rowcum[np.maximum(rows - 1, 0)], 0)
        cols = np.array([self._row(r)[k] for r, k in zip(rows, picks - before)], dtype=np.int64)
        return rows, cols

    @property
    def rows(self):
        return np.concatenate([np.full(len(self._row(r)), r, dtype=np.int64) for r in range(self.oh)]) if self.oh else np.zeros(0, np.int64)

    @property
    def cols(self):
        return np.concatenate([self._row(r) for r in range(self.oh)]) if self.oh else np.zeros(0, np.int64)


def _device_offset_table(M):
    """random_patch_offset_table for a mask that already lives on the device (the staged `self.M`): the 17 x 17 window-any runs as a
    kernel and (H - 14) x (W - 14) / 8 bytes + the row counts come back, instead of the whole mask going to the host for cumulative sums
    (17 ms per `set_input` at 1536 x 1536 with a real object mask)."""
    from . import _lib as L
    m = M[0, 0].contiguous().float()
    H, W = m.shape
    k, pad = 17, 1
    oh, ow = H + 2 * pad - k + 1, W + 2 * pad - k + 1
    words = (ow + 31) // 32
    scratch = torch.empty(H * ow, dtype=torch.uint8, device=m.device)
    bits = torch.empty((oh, words), dtype=torch.int32, device=m.device)
    rowcount = torch.empty(oh, dtype=torch.int32, device=m.device)
    L.call("skit_mask_box_bits", L.ptr(m), H, W, k, pad, L.ptr(scratch), L.ptr(bits), L.ptr(rowcount), L.stream())
    return _BitOffsetTable(bits.cpu().numpy().view(np.uint8).reshape(oh, words * 4), rowcount.cpu().numpy(), oh, ow)


def offset_table_arrays(m2d_device):
    """(bits uint8 [oh, 4 * words], rowcount int32 [oh]) host arrays of the candidate table for a device mask [H, W] (any dtype,
    > 0 = inside): what the device dataset stores with each item so that `set_input` needs no device work for it."""
    t = _device_offset_table(m2d_device[None, None])
    return t.bits, np.diff(np.concatenate([[0], t.rowcum])).astype(np.int32)


def offset_table_from_arrays(bits, rowcount, H, W):
    oh, ow = H + 2 - 17 + 1, W + 2 - 17 + 1
    return _BitOffsetTable(np.ascontiguousarray(bits).reshape(oh, -1), np.asarray(rowcount).reshape(oh), oh, ow)


def random_patch_offset_table(M):
    """`clamp(conv2d(M, ones(1,1,17,17), padding=1), 0, 1)` then torch.nonzero in row-major order
    (model_utils.py:212-218); the map is (H-14)x(W-14) and its (row, col) are used directly as
    (offset_y, offset_x).  Host side: for a non-negative mask "box sum > 0" is a 17x17 dilation, done separably with
    two integer running sums (exact for the 0/1 masks the datasets produce)."""
    if torch.is_tensor(M) and M.is_cuda:
        return _device_offset_table(M)
    m = np.asarray(M[0, 0].cpu()) > 0
    H, W = m.shape
    oh, ow = H + 2 - 17 + 1, W + 2 - 17 + 1
    if m.all():
        return _OffsetTable(None, oh, ow)
    p = np.zeros((H + 2, W + 3), dtype=np.int32)          # 1 zero halo (+ a leading column for the running sum)
    p[1:H + 1, 2:W + 2] = m
    cx = p.cumsum(1, dtype=np.int32)
    rowbox = (cx[:, 17:17 + ow] - cx[:, 0:ow]) > 0        # any mask pixel in the 17-wide window of each row
    q = np.zeros((H + 3, ow), dtype=np.int32)
    q[1:, :] = rowbox
    cy = q.cumsum(0, dtype=np.int32)
    return _OffsetTable((cy[17:17 + oh, :] - cy[0:oh, :]) > 0, oh, ow)


def get_patch_in_input(input, coords=None, sample_size=None, scale_multiplier=1, patch_size=32,
                       offset_x=None, offset_y=None, M=None, return_offset=False, **unused):
    """get_patch_in_input (model_utils.py:72-405) for the usages on the hot path: known coords;
    random offsets inside the mask; caller-supplied offsets.  input: [1, C, H, W] CUDA tensor."""
    if not input.is_cuda:
        raise RuntimeError("get_patch_in_input (B200 path) needs a CUDA tensor; there is no CPU fallback")
    patch_size = patch_size * scale_multiplier
    if coords is not None:
        cnp = coords.cpu().numpy() if torch.is_tensor(coords) else np.asarray(coords)
        assert cnp.shape[0] == 1, "coords should have batch size of 1"
        ox, oy, cs = find_coords_for_patch(cnp, scale_multiplier)
    else:
        assert sample_size is not None
        if offset_x is None:
            ox, oy = random_patch_offset_table(M).sample(sample_size)
        else:
            ox = np.asarray(offset_x.cpu() if torch.is_tensor(offset_x) else offset_x).reshape(-1).astype(np.int32)
            oy = np.asarray(offset_y.cpu() if torch.is_tensor(offset_y) else offset_y).reshape(-1).astype(np.int32)
        cs = np.full((sample_size,), patch_size, dtype=np.int32)
    if int(cs.max()) != patch_size:
        raise NotImplementedError("cutout != patch size (bicubic patch resize) is outside the hot path (T_resolution_multiplier = 1, resize_ratio = 1)")
    dev = input.device
    out = ops.patch_gather([input.contiguous().float()], torch.from_numpy(ox).to(dev), torch.from_numpy(oy).to(dev), int(patch_size))
    if return_offset:
        return out, ox / scale_multiplier, oy / scale_multiplier, cs / scale_multiplier
    return out


def compute_normal(T, scale_nz=0.25):
    """compute_normal (model_utils.py:418-425) — on the hot path this is fused into the generator head
    kernel (skit_g_head_fwd); this standalone form serves callers that hold only T (logging)."""
    n = torch.cat([T[:, 0:1], T[:, 1:2], torch.full_like(T[:, 0:1], scale_nz)], dim=1)
    return n / n.norm(dim=1, keepdim=True).clamp_min(1e-12)


def spe_grid(h, w, emb_dim=4, n=1):
    """SinusoidalPositionalEmbedding(emb_dim, padding_idx=0).make_grid2d(h, w, n)
    (positional_encoding.py:61-86,113-159): a constant of (h, w); built once on the host and cached on
    the device by the model.  Channels: sin/cos of the x position (1..w), then of the y position."""
    half = emb_dim // 2
    freq = torch.exp(torch.arange(half, dtype=torch.float32) * -(math.log(10000.0) / (half - 1)))

    def table(length):
        pos = torch.arange(1, length + 1, dtype=torch.float32)[:, None] * freq[None, :]
        return torch.cat([torch.sin(pos), torch.cos(pos)], dim=1)

    ex = table(w).t()[None, :, None, :].expand(n, emb_dim, h, w)
    ey = table(h).t()[None, :, :, None].expand(n, emb_dim, h, w)
    return torch.cat([ex, ey], dim=1).contiguous()
