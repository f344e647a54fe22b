"""TEST INFRASTRUCTURE ONLY (oracle/): CPU restatement of the arithmetic behind the reference's data pipeline
(`data/singleskit_dataset.py`, `data/dataset_util.py`).  numpy, integer / fp64 exact.

The resampling itself lives in a third-party dependency the reference does not vendor: **Pillow** (`requirements.txt`,
unpinned; this container has 12.2.0) — `Image.resize(size, Image.LANCZOS)` called from `dataset_util.py:159-163`
(`zoom_img`), `:194-201` (`crop_img`), `:219-231` (`make_power_2_img`).  `pil_resize_u8` restates Pillow's published
8-bit two-pass algorithm (`src/libImaging/Resample.c`: `precompute_coeffs`, `normalize_coeffs_8bpc`,
`ImagingResampleHorizontal_8bpc` / `Vertical_8bpc`); tests pin it bit-exactly against the installed Pillow.
`laplacian_var_u8` restates `util/util.py:261-265` (`cv2.Laplacian(image - ref, CV_64F).var()`, ksize 1, reflect-101 border,
uint8 wrap-around of `image - ref`).  Parity status: pinned (Pillow / OpenCV themselves in tests/test_data_oracle.py, the
reference's own `SingleSkitDataset` output in tests/golden/data_pipeline.npz).
"""
import math

import numpy as np

BOX, BILINEAR, HAMMING, BICUBIC, LANCZOS = 4, 2, 5, 3, 1       # PIL.Image.Resampling values
PRECISION_BITS = 32 - 8 - 2


def _sinc(x):
    if x == 0.0:
        return 1.0
    x = x * math.pi
    return math.sin(x) / x


def _filter(kind):
    if kind == LANCZOS:
        return 3.0, lambda x: _sinc(x) * _sinc(x / 3) if -3.0 <= x < 3.0 else 0.0
    if kind == BICUBIC:
        a = -0.5

        def f(x):
            x = abs(x)
            if x < 1.0:
                return ((a + 2.0) * x - (a + 3.0)) * x * x + 1
            if x < 2.0:
                return (((x - 5) * x + 8) * x - 4) * a
            return 0.0
        return 2.0, f
    if kind == BILINEAR:
        return 1.0, lambda x: 1.0 - abs(x) if abs(x) < 1.0 else 0.0
    if kind == BOX:
        return 0.5, lambda x: 1.0 if -0.5 < x <= 0.5 else 0.0
    if kind == HAMMING:
        def h(x):
            x = abs(x)
            if x == 0.0:
                return 1.0
            if x >= 1.0:
                return 0.0
            x = x * math.pi
            return math.sin(x) / x * (0.54 + 0.46 * math.cos(x))
        return 1.0, h
    raise ValueError("unsupported filter %r" % (kind,))


def precompute_coeffs(in_size, out_size, kind):
    """Resample.c `precompute_coeffs` + `normalize_coeffs_8bpc` for the full box: (ksize, bounds[out,2], kk[out,ksize] int32)."""
    support0, f = _filter(kind)
    scale = filterscale = in_size / out_size
    if filterscale < 1.0:
        filterscale = 1.0
    support = support0 * filterscale
    ksize = int(math.ceil(support)) * 2 + 1
    bounds = np.zeros((out_size, 2), np.int32)
    kk = np.zeros((out_size, ksize), np.int32)
    ss = 1.0 / filterscale
    for xx in range(out_size):
        center = (xx + 0.5) * scale
        xmin = int(center - support + 0.5)
        if xmin < 0:
            xmin = 0
        xmax = int(center + support + 0.5)
        if xmax > in_size:
            xmax = in_size
        xmax -= xmin
        w = [f((x + xmin - center + 0.5) * ss) for x in range(xmax)]
        ww = 0.0
        for v in w:
            ww += v
        for x in range(xmax):
            v = w[x] / ww if ww != 0.0 else w[x]
            kk[xx, x] = int(-0.5 + v * (1 << PRECISION_BITS)) if v < 0 else int(0.5 + v * (1 << PRECISION_BITS))
        bounds[xx] = (xmin, xmax)
    return ksize, bounds, kk


def _pass(src, out_size, kind):
    """One 8-bit pass along axis 1 of src [rows, in, c] -> [rows, out, c]."""
    rows, n, c = src.shape
    ksize, bounds, kk = precompute_coeffs(n, out_size, kind)
    idx = np.minimum(bounds[:, :1] + np.arange(ksize)[None, :], n - 1)          # [out, ksize]; taps beyond xmax have coefficient 0
    acc = np.full((rows, out_size, c), 1 << (PRECISION_BITS - 1), np.int64)
    for k in range(ksize):
        acc += src[:, idx[:, k], :].astype(np.int64) * kk[:, k].astype(np.int64)[None, :, None]
    return np.clip(acc >> PRECISION_BITS, 0, 255).astype(np.uint8)


def pil_resize_u8(img, out_h, out_w, kind=LANCZOS):
    """`PIL.Image.resize((out_w, out_h), kind)` for an 8-bit image [h, w] or [h, w, c]: horizontal pass, then vertical."""
    a = img[:, :, None] if img.ndim == 2 else img
    h, w, _ = a.shape
    if (h, w) == (out_h, out_w):
        return img.copy()
    if w != out_w:
        a = _pass(a, out_w, kind)
    if h != out_h:
        a = _pass(a.transpose(1, 0, 2), out_h, kind).transpose(1, 0, 2)
    a = np.ascontiguousarray(a)
    return a[:, :, 0] if img.ndim == 2 else a


def to_tensor_norm(img_u8, normalize=True):
    """`transforms.ToTensor()` [+ `Normalize(0.5, 0.5)`] (`singleskit_dataset.py:301-315`): fp32 CHW."""
    a = img_u8[:, :, None] if img_u8.ndim == 2 else img_u8
    t = a.transpose(2, 0, 1).astype(np.float32) / np.float32(255)
    if normalize:
        t = (t - np.float32(0.5)) / np.float32(0.5)
    return t


def crop_zero(img, x0, y0, size):
    """`PIL.Image.crop((x0, y0, x0+size, y0+size))`: pixels outside the image read 0."""
    out = np.zeros((size, size) + img.shape[2:], img.dtype)
    h, w = img.shape[:2]
    ys, xs = max(0, y0), max(0, x0)
    ye, xe = min(h, y0 + size), min(w, x0 + size)
    if ye > ys and xe > xs:
        out[ys - y0:ye - y0, xs - x0:xe - x0] = img[ys:ye, xs:xe]
    return out


def contact_centers(touch_mask, center_mask, M3, roi_x, roi_y, patch=32):
    """The centre loop of `process_all_valid_patches` (`singleskit_dataset.py:768-803`, T_resolution_multiplier = 1):
    row-major list of the (cx, cy) with `center_mask > 0` whose patch x patch window of `touch_mask * M_patch / 255` reaches 1."""
    out = []
    half = patch // 2
    ys, xs = np.where(center_mask > 0)
    for cx, cy in zip(xs, ys):
        sq = touch_mask[cy - half:cy + half, cx - half:cx + half]
        mp = crop_zero(M3, int(np.round(roi_x + (cx - half))), int(np.round(roi_y + (cy - half))), patch)
        sq = sq * mp / 255
        if np.max(sq) >= 1:
            out.append((int(cx), int(cy)))
    return out


def laplacian_var_u8(patch_u8, ref=255):
    """`variance_of_laplacian(S_patch, ref=ones*255)` (`util/util.py:261-265`): uint8 wrap of (image - ref), 3x3 Laplacian with
    reflect-101 borders in fp64, population variance."""
    a = (patch_u8.astype(np.int64) - ref) % 256
    p = np.pad(a, 1, mode="reflect")
    lap = p[:-2, 1:-1] + p[2:, 1:-1] + p[1:-1, :-2] + p[1:-1, 2:] - 4 * a
    return float(np.var(lap.astype(np.float64)))
