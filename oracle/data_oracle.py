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


# ------------------------------------------------------------------------------------------------------------------------------
# The dataset itself: a plain numpy restatement of SingleSkitDataset.preprocess_data (data/singleskit_dataset.py:194-432),
# find_validate_touch_patches_and_coords (:434-660) and process_all_valid_patches (:662-1128) for the cases the reference's own
# data satisfies (every touch patch inside every crop, contact masks present, T_resolution_multiplier 1).  Draws from Python's
# `random` / `np.random` in the reference's call order, so a seeded run reproduces its items (tests/golden/data_pipeline.npz).
# Also the CPU arm of tools/bench_data.py.
def _walk_files(d, suffix):
    import os
    out = []
    for root, _, fnames in sorted(os.walk(d, followlinks=True)):            # data/image_folder.py:28-61
        for f in fnames:
            if f.endswith(suffix):
                out.append(os.path.join(root, f))
    return out


def _load_touch(path):
    z = np.load(path)                                                       # data/dataset_util.py:5-62
    tm, cm = z["touch_thresh"], z["touch_center_thresh"]
    if np.max(tm) > 1:
        tm = tm / 255
    if np.max(cm) > 1:
        cm = cm / 255
    return z["gx_raw"], z["gy_raw"], z["vision_mask_x"], z["vision_mask_y"], z["vision_mask_h"], z["vision_mask_w"], tm, cm


def dataset_items(opt):
    """-> list of item dicts (numpy arrays; S / I / M as the final uint8 images `S_u8` ...)."""
    import os
    import random
    from PIL import Image
    n_items = opt.data_len
    S = np.array(Image.open(_walk_files(os.path.join(opt.dataroot, opt.subdir_S), ".png")[0]).convert("L"))
    I = np.array(Image.open(_walk_files(os.path.join(opt.dataroot, opt.subdir_I), ".png")[0]).convert("RGB"))
    M = np.array(Image.open(_walk_files(os.path.join(opt.dataroot, opt.subdir_M), ".png")[0]).convert("L"))
    sets = {"": [_load_touch(p) for p in _walk_files(os.path.join(opt.dataroot, opt.subdir_T), "_tactile.npz")]}
    if opt.subdir_valT is not None:
        sets["val_"] = [_load_touch(p) for p in _walk_files(os.path.join(opt.dataroot, opt.subdir_valT), "_tactile.npz")]
    A_zoom = 1 / opt.random_scale_max if opt.is_train else 1                # :176-186
    zoom = np.random.uniform(A_zoom, 1.0, size=(n_items // opt.batch_size + 1, 1, 2))
    zoom = np.reshape(np.tile(zoom, (1, opt.batch_size, 1)), [-1, 2])
    items = []
    crop = opt.crop_size
    for index in range(n_items):
        sf_h, sf_w = (zoom[0] if "zoom" in opt.preprocess else (1, 1))      # :238-254 (always level 0)

        def stage12(img):
            h, w = img.shape[:2]
            if "zoom" in opt.preprocess:
                img = pil_resize_u8(img, int(round(h * sf_h)), int(round(w * sf_w)))        # dataset_util.py:159-163
            h, w = img.shape[:2]
            ratio = 1 if (w >= crop and h >= crop) else max(crop / w, crop / h)             # :188-192
            return pil_resize_u8(img, int(round(h * ratio)), int(round(w * ratio))), ratio
        S2, ratio = stage12(S)
        I2, _ = stage12(I)
        M2, _ = stage12(M)
        h, w = S2.shape[:2]
        if "crop" not in opt.preprocess:                                     # dataset_util.py:165-183
            cx, cy = (w - crop) // 2, (h - crop) // 2
        elif opt.center_w > 0 or opt.center_h > 0:
            buf = min(max(0, (w - opt.center_w) // 2), max(0, (h - opt.center_h) // 2), h - crop, w - crop)
            cx = random.randint(0, buf); cy = random.randint(0, buf)
        else:
            cx = random.randint(0, max(0, w - crop)); cy = random.randint(0, max(0, h - crop))
        t = int(round(crop / 256) * 256)                                     # :219-231
        rr = 1 if t == crop else t / crop
        fin = lambda a: pil_resize_u8(a[cy:cy + crop, cx:cx + crop], t, t)
        S3, I3, M3 = fin(S2), fin(I2), fin(M2)
        item = {"S_u8": S3, "I_u8": I3, "M_u8": M3,
                "augmentation_params": {"H": S.shape[1], "W": S.shape[0], "scale_factor_h": sf_h, "scale_factor_w": sf_w, "crop_size_h": crop,
                                        "crop_size_w": crop, "resize_ratio": ratio, "crop_pos_x": cx, "crop_pos_y": cy, "resize_ratio_w": rr,
                                        "resize_ratio_h": rr, "patch_crop_size": 32}}
        for pre, patches in sets.items():
            rois = []
            for (_, _, x, y, ph, pw, _, _) in patches:                       # :492-560
                if "padded" in opt.dataroot:
                    ps = int(opt.dataroot.split("padded_")[1].split("/")[0].split("_")[0])
                    x = x + (ps - opt.center_w) // 2; y = y + (ps - opt.center_h) // 2
                x1, y1, h1, w1 = x * sf_w, y * sf_h, ph * sf_h, pw * sf_w
                x2, y2, h2, w2 = x1 * ratio - cx, y1 * ratio - cy, h1 * ratio, w1 * ratio
                assert not (x2 < 0 or x2 + w2 > crop or y2 < 0 or y2 + h2 > crop), "patch outside the crop (singleskit_dataset.py:742-751)"
                rois.append([int(round(x2 * rr)), int(round(y2 * rr)), int(round(h2 * rr)), int(round(w2 * rr))])
            T_images, coords, masks, full = [], [], [], []
            for (gx, gy, _, _, _, _, tm, cm), (x3, y3, h3, w3) in zip(patches, rois):       # :738-860
                if np.sum(M3[y3:y3 + h3, x3:x3 + w3]) == 0:
                    continue
                full.append([x3, y3, h3, w3])
                cs = contact_centers(tm, cm, M3, x3, y3)
                num = min(len(cs), opt.sample_bbox_per_patch)
                sel = random.sample(range(len(cs)), num) if opt.is_train else np.arange(len(cs) // 2, len(cs) // 2 + num)
                for k in sel:
                    px, py = cs[k]
                    win = (slice(py - 16, py + 16), slice(px - 16, px + 16))
                    T_images.append(np.stack([gx[win], gy[win]]))
                    coords.append([x3, y3, h3, w3, 32, 1, px - 16, py - 16])
                    masks.append(tm[win] * crop_zero(M3, x3 + px - 16, y3 + py - 16, 32) / 255)
            n = len(coords)
            if pre == "" and opt.is_train and opt.w_resampling:              # :1000-1056, :623-631
                wts = [min(max(opt.resampling_w_min, laplacian_var_u8(crop_zero(S3, c[0] + c[6], c[1] + c[7], 32))), opt.resampling_w_max)
                       for c in coords]
                pick = random.choices(range(n), weights=np.array(wts), k=min(opt.batch_size_G2, n) if opt.batch_size_G2 > 0 else n)
            elif opt.is_train:
                k = opt.batch_size_G2 if pre == "" else opt.batch_size_G2_val
                pick = random.sample(range(n), min(k, n) if k > 0 else n)
            else:
                pick = list(range(n))
            item[pre + "T_images"] = np.stack(T_images)[pick]
            item[pre + "T_coords"] = np.stack(coords)[pick]
            item[pre + "I_masks"] = np.stack(masks)[pick]
            item[pre + "full_T_coords"] = np.array(full)
        items.append(item)
    return items
