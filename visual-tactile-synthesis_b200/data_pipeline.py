"""Data pipeline of the sinskitG model on the device: `SingleSkitDataset` with the reference's
constructor, options, `__getitem__` / `__len__` and item layout (data/singleskit_dataset.py:28-1147), reading the reference's
on-disk format (`trainS/ trainI/ trainM/` images, `trainT/ valT/` `*_tactile.npz`, data/dataset_util.py:5-62).

What the reference does on the host for every one of its `data_len` augmentations — Pillow LANCZOS zoom / crop / power-of-2
resize of the sketch, image and mask, then a Python loop over every centre pixel of every touch patch with a PIL crop inside
(20-30 min at start-up, README.md:129) — runs here as byte / integer kernels (csrc/data_ops.cu) on sources that stay resident in
HBM as uint8:

  * the zoomed / ratio-resized sources are computed ONCE (the reference recomputes the same zoom for every index:
    `self.zoom_levels_A[0]`, singleskit_dataset.py:241) with Pillow's exact fixed-point resampler;
  * per augmentation only the crop position and the touch-patch selection are drawn — through Python's `random` in the
    reference's own call order, so a seeded run returns the reference's items bit for bit (tests/golden/data_pipeline.npz);
  * the contact-centre search, the 32 x 32 gathers and the Laplacian-variance weights are batched over all patches;
  * full-resolution S / I / M tensors are produced on demand by one crop + ToTensor + Normalize kernel each (and cached while
    they fit `cache_bytes`), instead of holding `data_len` fp32 copies (9.4 GB at the defaults) in host memory.

Items hold CUDA tensors (`set_input` then copies nothing); `T_coords`, `full_T_coords` and `augmentation_params` are host objects
as in the reference.  No CPU fallback: without the CUDA library the constructor raises.
"""
import ntpath
import os
import random

import numpy as np
import torch

from . import _lib as L

_p = L.ptr

LANCZOS, BILINEAR, BICUBIC, BOX, HAMMING = 1, 2, 3, 4, 5       # PIL.Image.Resampling values

IMG_EXTENSIONS = ['.jpg', '.JPG', '.jpeg', '.JPEG', '.png', '.PNG', '.ppm', '.PPM', '.bmp', '.BMP', '.tif', '.TIF', '.tiff', '.TIFF']


# ------------------------------------------------------------------------------------------------ device image ops
def _hwc(img):
    if not (img.is_cuda and img.dtype == torch.uint8 and img.is_contiguous() and img.dim() == 3):
        raise RuntimeError("data pipeline (B200 path) works on contiguous CUDA uint8 [H, W, C] images; there is no CPU fallback")
    return img


def resize_u8(img, out_h, out_w, method=LANCZOS):
    """`PIL.Image.resize((out_w, out_h), method)` on a device image [H, W, C] uint8, bit-identical with Pillow."""
    _hwc(img)
    h, w, c = img.shape
    out = torch.empty((out_h, out_w, c), dtype=torch.uint8, device=img.device)
    L.call("skit_resize_u8", _p(img), h, w, c, _p(out), int(out_h), int(out_w), int(method), L.stream())
    return out


def crop_to_tensor(img, x0, y0, w, h, normalize=True):
    """`transforms.ToTensor()(img.crop((x0, y0, x0+w, y0+h)))` [+ Normalize(0.5, 0.5)]: fp32 [C, h, w]."""
    _hwc(img)
    sh, sw, c = img.shape
    out = torch.empty((c, h, w), dtype=torch.float32, device=img.device)
    L.call("skit_u8_crop_to_tensor", _p(img), sh, sw, c, int(y0), int(x0), int(h), int(w), int(bool(normalize)), _p(out), L.stream())
    return out


def size_of(img):
    """PIL's `img.size` = (width, height)."""
    return (img.shape[1], img.shape[0])


# ------------------------------------------------------------------------------------------------ dataset_util.py mirrors
def zoom_img(img, scale_factor_h=1, scale_factor_w=1, method=BICUBIC):
    """dataset_util.py:159-163."""
    ow, oh = size_of(img)
    nw, nh = ow * scale_factor_w, oh * scale_factor_h
    return resize_u8(img, int(round(nh)), int(round(nw)), method)


def zoom_find_coords(ROI_x, ROI_y, ROI_h, ROI_w, scale_factor_h=1, scale_factor_w=1):
    """dataset_util.py:152-157."""
    return ROI_x * scale_factor_w, ROI_y * scale_factor_h, ROI_h * scale_factor_h, ROI_w * scale_factor_w


def get_params(size, crop_size_h=512, crop_size_w=512, center_w=0, center_h=0, center_crop=False):
    """dataset_util.py:165-183: the crop origin; draws from Python's `random` exactly as the reference does."""
    w, h = size
    assert w >= crop_size_w and h >= crop_size_h, "The image is smaller than crop_size. Cannot perform get_params for cropping"
    assert crop_size_h >= center_h and crop_size_w >= center_w, "crop_size h {} w {} cannot cover the center region h {} w {}".format(
        crop_size_h, crop_size_w, center_h, center_w)
    if center_crop:
        x = (w - crop_size_w) // 2
        y = (h - crop_size_h) // 2
    elif center_w > 0 or center_h > 0:
        buffer = min(np.maximum(0, (w - center_w) // 2), np.maximum(0, (h - center_h) // 2), h - crop_size_h, w - crop_size_w)
        x = random.randint(0, buffer)
        y = random.randint(0, buffer)
    else:
        x = random.randint(0, np.maximum(0, w - crop_size_w))
        y = random.randint(0, np.maximum(0, h - crop_size_h))
    return (x, y)


def crop_resize_ratio(size, crop_size_h, crop_size_w):
    """The ratio `crop_img` derives when none is given (dataset_util.py:188-192)."""
    w, h = size
    if w >= crop_size_w and h >= crop_size_h:
        return 1
    return max(crop_size_w / w, crop_size_h / h)


def crop_find_coords(ROI_x, ROI_y, ROI_h, ROI_w, crop_size_h, crop_size_w, resize_ratio, crop_pos_x, crop_pos_y):
    """dataset_util.py:204-217."""
    ROI_x = ROI_x * resize_ratio; ROI_y = ROI_y * resize_ratio
    ROI_h = ROI_h * resize_ratio; ROI_w = ROI_w * resize_ratio
    new_ROI_x = ROI_x - crop_pos_x
    new_ROI_y = ROI_y - crop_pos_y
    if new_ROI_x < 0 or new_ROI_x + ROI_w > crop_size_w:
        return False, new_ROI_x, new_ROI_y, ROI_h, ROI_w
    elif new_ROI_y < 0 or new_ROI_y + ROI_h > crop_size_h:
        return False, new_ROI_x, new_ROI_y, ROI_h, ROI_w
    return True, new_ROI_x, new_ROI_y, ROI_h, ROI_w


def make_power_2_size(size, base):
    """The target size and ratios of `make_power_2_img` (dataset_util.py:219-231): (w, h, resize_ratio_w, resize_ratio_h)."""
    ow, oh = size
    h = int(round(oh / base) * base)
    w = int(round(ow / base) * base)
    if h == oh and w == ow:
        return ow, oh, 1, 1
    return w, h, w / ow, h / oh


def make_power_2_find_coords(ROI_x, ROI_y, ROI_h, ROI_w, resize_ratio_w, resize_ratio_h):
    """dataset_util.py:233-238."""
    return ROI_x * resize_ratio_w, ROI_y * resize_ratio_h, ROI_h * resize_ratio_h, ROI_w * resize_ratio_w


def global_padding_find_coords(ROI_x, ROI_y, ROI_h, ROI_w, org_w=1280, org_h=960, padded_size=1600):
    """dataset_util.py:240-243."""
    return ROI_x + (padded_size - org_w) // 2, ROI_y + (padded_size - org_h) // 2, ROI_h, ROI_w


def touch_data_loader(path, return_mask=True):
    """dataset_util.py:5-62 with `convert2im=False`: (gx, gy, ROI_x, ROI_y, ROI_h, ROI_w, touch_mask, touch_center_mask)."""
    npz_data = np.load(path)
    ROI_x, ROI_y = npz_data["vision_mask_x"], npz_data["vision_mask_y"]
    ROI_h, ROI_w = npz_data["vision_mask_h"], npz_data["vision_mask_w"]
    gx, gy = npz_data["gx_raw"], npz_data["gy_raw"]
    touch_mask = touch_center_mask = None
    if return_mask:
        assert 'touch_thresh' in npz_data.files, "touch_thresh not found in npz_data"
        assert 'touch_center_thresh' in npz_data.files, "touch_center_thresh not found in npz_data"
        touch_mask = npz_data["touch_thresh"]; touch_center_mask = npz_data["touch_center_thresh"]
        if np.max(touch_mask) > 1:
            touch_mask = touch_mask / 255
        if np.max(touch_center_mask) > 1:
            touch_center_mask = touch_center_mask / 255
    return gx, gy, ROI_x, ROI_y, ROI_h, ROI_w, touch_mask, touch_center_mask


def is_image_file(filename):
    return any(filename.endswith(extension) for extension in IMG_EXTENSIONS)


def make_dataset(dir, max_dataset_size=float("inf")):
    """data/image_folder.py:28-37."""
    images = []
    assert os.path.isdir(dir) or os.path.islink(dir), '%s is not a valid directory' % dir
    for root, _, fnames in sorted(os.walk(dir, followlinks=True)):
        for fname in fnames:
            if is_image_file(fname):
                images.append(os.path.join(root, fname))
    return images[:min(max_dataset_size, len(images))]


def make_touch_image_dataset(dir, max_dataset_size=float("inf")):
    """data/image_folder.py:40-61 (same walk order as the reference: directories sorted, file names as the OS lists them)."""
    assert os.path.isdir(dir) or os.path.islink(dir), '%s is not a valid directory for tactile image dataset' % dir
    if len(os.listdir(dir)) == 0:
        print("Empty directory for %s, return empty list for touch data" % (dir))
        return [], []
    paths = []
    for root, _, fnames in sorted(os.walk(dir, followlinks=True)):
        for fname in fnames:
            if fname.endswith('_tactile.npz'):
                paths.append(os.path.join(root, fname))
    return paths[:min(max_dataset_size, len(paths))]


def load_image_u8(path, mode, device):
    """`Image.open(path)` -> grayscale (`ImageOps.grayscale`) or RGB, as a device uint8 [H, W, C] image.  Decoding stays on the
    host (Pillow, once per file); everything after it runs on the device."""
    from PIL import Image
    img = Image.open(path).convert(mode)
    a = np.asarray(img)
    if a.ndim == 2:
        a = a[:, :, None]
    return torch.from_numpy(np.array(a, copy=True)).to(device)


# ------------------------------------------------------------------------------------------------ touch patches on the device
class TouchSet:
    """All `*_tactile.npz` files of one directory, back to back in HBM: gx / gy in the files' own dtype, the contact mask in fp64
    (as the reference holds it after its /255), the centre mask as bytes, plus the per-patch geometry."""

    def __init__(self, paths, device, patch=32):
        self.paths = list(paths)
        self.P = len(self.paths)
        self.device = device
        self.patch = patch
        gxs, gys, tms, cms, self.roi, self.hw = [], [], [], [], [], []
        half = patch // 2
        for path in self.paths:
            gx, gy, ROI_x, ROI_y, ROI_h, ROI_w, tm, cm = touch_data_loader(path)
            assert tm is not None and cm is not None, "Need valid touch mask and touch center mask"
            if gx.shape != gy.shape or gx.shape != tm.shape or gx.shape != cm.shape:
                raise ValueError("%s: gx / gy / touch_thresh / touch_center_thresh shapes differ" % path)
            ys, xs = np.where(cm > 0)
            if len(ys) and (ys.min() < half or xs.min() < half or ys.max() > gx.shape[0] - half or xs.max() > gx.shape[1] - half):
                # the reference slices touch_mask[cy-16:cy+16, cx-16:cx+16]; a centre closer than 16 pixels to the border gives a
                # short (or, for a negative start, empty) window there and the product with the 32 x 32 mask patch raises
                raise ValueError("%s: a contact-centre pixel lies within %d pixels of the patch border" % (path, half))
            gxs.append(np.ascontiguousarray(gx)); gys.append(np.ascontiguousarray(gy))
            tms.append(np.asarray(tm, np.float64)); cms.append((cm > 0).astype(np.uint8))
            self.roi.append((ROI_x, ROI_y, ROI_h, ROI_w))
            self.hw.append(gx.shape)
        if self.P == 0:
            return
        dt = gxs[0].dtype
        if any(a.dtype != dt for a in gxs + gys) or dt.itemsize not in (4, 8):
            raise ValueError("gx_raw / gy_raw must share one 4- or 8-byte dtype across the touch files")
        self.np_dtype = dt
        self.torch_dtype = torch.from_numpy(np.zeros(1, dt)).dtype
        off = np.zeros(self.P + 1, np.int64)
        off[1:] = np.cumsum([h * w for h, w in self.hw])
        self.total = int(off[-1])
        self.pix_off_host = off
        cat = lambda arrs: torch.from_numpy(np.concatenate([a.reshape(-1) for a in arrs])).to(device)
        self.gx, self.gy, self.tm, self.cm = cat(gxs), cat(gys), cat(tms), cat(cms)
        self.pix_off = torch.from_numpy(off).to(device)
        self.ph = torch.tensor([h for h, _ in self.hw], dtype=torch.int32, device=device)
        self.pw = torch.tensor([w for _, w in self.hw], dtype=torch.int32, device=device)
        self.scratch = torch.empty(3 * self.total, dtype=torch.uint8, device=device)
        self.centers = torch.empty(self.total, dtype=torch.int32, device=device)
        self.counts = torch.empty(self.P, dtype=torch.int32, device=device)
        self.in_mask = torch.empty(self.P, dtype=torch.int32, device=device)

    def contact_centers(self, M3, roi_x, roi_y):
        """-> host (counts[P], in_mask[P]); the ordered centre lists stay on the device in `self.centers`."""
        mh, mw, _ = M3.shape
        self.roi_x = torch.tensor(roi_x, dtype=torch.int32, device=self.device)
        self.roi_y = torch.tensor(roi_y, dtype=torch.int32, device=self.device)
        L.call("skit_contact_centers", _p(self.tm), _p(self.cm), _p(self.pix_off), self.total, _p(self.ph), _p(self.pw), _p(self.roi_x),
               _p(self.roi_y), self.P, _p(M3), mh, mw, self.patch, _p(self.scratch), _p(self.counts), _p(self.centers), _p(self.in_mask),
               L.stream())
        both = torch.stack([self.counts, self.in_mask]).cpu().numpy()
        return both[0], both[1]

    def squares(self, M3, sel_patch, sel_rank):
        """For selections (patch index, rank within that patch's ordered centre list): (cx, cy) on the host and the device tensors
        T_images [K, 2, patch, patch], I_masks [K, patch, patch] fp64."""
        K = len(sel_patch)
        mh, mw, _ = M3.shape
        sp = torch.tensor(sel_patch, dtype=torch.int64, device=self.device)
        lin = self.centers[self.pix_off[sp] + torch.tensor(sel_rank, dtype=torch.int64, device=self.device)]
        w = self.pw[sp]
        cx = (lin % w).to(torch.int32).contiguous()
        cy = (lin // w).to(torch.int32).contiguous()
        sp32 = sp.to(torch.int32).contiguous()
        T = torch.empty((K, 2, self.patch, self.patch), dtype=self.torch_dtype, device=self.device)
        Mk = torch.empty((K, self.patch, self.patch), dtype=torch.float64, device=self.device)
        L.call("skit_touch_squares", _p(self.tm), _p(self.pix_off), _p(self.ph), _p(self.pw), _p(self.roi_x), _p(self.roi_y), self.P, _p(M3),
               mh, mw, _p(self.gx), _p(self.gy), self.np_dtype.itemsize, _p(sp32), _p(cx), _p(cy), K, self.patch, _p(T), _p(Mk), L.stream())
        cxy = torch.stack([cx, cy]).cpu().numpy()
        return cxy[0], cxy[1], T, Mk


def laplacian_var(img1, x0, y0, size, ref=255):
    """`variance_of_laplacian(S3.crop(...), ref=255)` for K windows of a single-band device image -> host fp64 [K]."""
    _hwc(img1)
    h, w, c = img1.shape
    assert c == 1, "the resampling weight is defined on the 1-channel sketch (util/util.py:261-265 on a 2-D patch)"
    K = len(x0)
    xs = torch.tensor(x0, dtype=torch.int32, device=img1.device)
    ys = torch.tensor(y0, dtype=torch.int32, device=img1.device)
    out = torch.empty(K, dtype=torch.float64, device=img1.device)
    L.call("skit_laplacian_var_u8", _p(img1), h, w, _p(xs), _p(ys), K, int(size), int(ref), _p(out), L.stream())
    return out.cpu().numpy()


# ------------------------------------------------------------------------------------------------ the dataset
class _Material:
    """One object's resident sources: decoded uint8 images and its two touch sets."""

    def __init__(self, S_path, S_img, I_img, M_img, M_path, touch, val_touch, extra=None):
        self.S_path, self.S_img, self.I_img, self.M_img, self.M_path = S_path, S_img, I_img, M_img, M_path
        self.touch, self.val_touch = touch, val_touch
        self.extra = extra or {}          # further images that ride through the same zoom / crop / resize chain (skitG's style_I / style_M)


def _str2bool(v):        # util/util.py str2bool
    if isinstance(v, bool):
        return v
    if v.lower() in ("yes", "true", "t", "y", "1"):
        return True
    if v.lower() in ("no", "false", "f", "n", "0"):
        return False
    raise ValueError("Boolean value expected.")


class SingleSkitDataset(torch.utils.data.Dataset):
    """data/singleskit_dataset.py:28 — one sketch / mask / image, `data_len` augmentations, touch patches with their coordinates."""

    @staticmethod
    def modify_commandline_options(parser, is_train):
        """singleskit_dataset.py:43-82 (skit_dataset.py:43-84 declares the same arguments)."""
        parser.add_argument("--subdir_S", type=str, default="trainS", help="subdirectory for S input")
        parser.add_argument("--subdir_I", type=str, default="trainI", help="subdirectory for I input")
        parser.add_argument("--subdir_T", type=str, default="trainT", help="subdirectory for T input")
        parser.add_argument("--subdir_M", type=str, default="trainM", help="subdirectory for mask input")
        parser.add_argument("--subdir_valT", type=str, default="valT", help="subdirectory for T input for validation")
        parser.add_argument("--is_train", type=_str2bool, default=True, help="whether the model is in training mode")
        if is_train:
            parser.set_defaults(subdir_S="trainS", subdir_I="trainI", subdir_T="trainT", subdir_M="trainM", subdir_valT="valT", is_train=True)
        else:
            parser.set_defaults(subdir_S="testS", subdir_I="testI", subdir_T="testT", subdir_M="testM", subdir_valT=None, is_train=False)
        return parser

    PATCH = 32

    def __init__(self, opt, verbose=False, default_len=1000, device=None, cache_bytes=8 << 30, host_items=None):
        """`host_items` (or `opt.data_host_items`): return host tensors, as the reference's items are — for its own
        `CustomDatasetDataLoader` (data/__init__.py:75-82), whose `pin_memory=True` cannot take CUDA tensors.  Default: device items."""
        L.load()     # fails loudly when the CUDA library is missing
        if not torch.cuda.is_available():
            raise RuntimeError("%s (B200 path) needs a CUDA device; there is no CPU fallback" % type(self).__name__)
        self.opt = opt
        self.root = opt.dataroot
        self.current_epoch = 0
        self.verbose = verbose
        self.device = torch.device(device if device is not None else "cuda:%d" % torch.cuda.current_device())
        self.data_dict = {}
        self.data_len = opt.data_len if hasattr(opt, "data_len") else default_len
        self._cache_budget = cache_bytes
        self.host_items = bool(getattr(opt, "data_host_items", False)) if host_items is None else bool(host_items)
        self._cache_used = 0
        self._image_cache = {}
        self._host_cache = {}
        self._src_cache = {}
        self._index_key = {}
        if getattr(opt, "T_resolution_multiplier", 1) != 1:
            raise NotImplementedError("T_resolution_multiplier != 1 is outside the B200 path (DESIGN.md section 10)")
        self._load_materials()
        for m in self.materials:
            if m.touch is not None and m.M_img is None:
                raise ValueError("touch patches are validated against the object mask (singleskit_dataset.py:742-745): use_bg_mask must be True")
        A_zoom = 1 / self.opt.random_scale_max if self.opt.is_train else 1
        zoom_levels_A = np.random.uniform(A_zoom, 1.0, size=(len(self) // opt.batch_size + 1, 1, 2))
        self.zoom_levels_A = np.reshape(np.tile(zoom_levels_A, (1, opt.batch_size, 1)), [-1, 2])
        self.preprocess_data()

    # ---- what differs between the single- and the multi-material dataset
    def _material_index(self, index):
        return 0

    def _zoom_index(self, index):
        return 0                # singleskit_dataset.py:241 reads zoom_levels_A[0] for every index

    def _name(self, m):
        return os.path.splitext(ntpath.basename(m.S_path[0]))[0]      # `S_path[0]` of a str: the reference's own quirk (:409-410)

    def _load_sketch(self, path):
        opt = self.opt
        if opt.sketch_nc == 1:
            return load_image_u8(path, "L", self.device)
        assert opt.sketch_nc == 3, "Load sketch either in grayscale or RGB"
        return load_image_u8(path, "RGB", self.device)

    def _load_materials(self):
        """singleskit_dataset.py:84-173."""
        opt = self.opt
        self.dir_S = os.path.join(opt.dataroot, opt.subdir_S)
        self.dir_I = os.path.join(opt.dataroot, opt.subdir_I)
        self.dir_T = os.path.join(opt.dataroot, opt.subdir_T)
        self.dir_M = os.path.join(opt.dataroot, opt.subdir_M)
        self.is_train = opt.is_train
        if opt.subdir_valT is not None:
            self.dir_valT = os.path.join(opt.dataroot, opt.subdir_valT)
            assert os.path.exists(self.dir_valT), "missing val T data for train datasets {}".format(self.dir_valT)
        assert os.path.exists(self.dir_S), "missing S data for datasets {}".format(self.dir_S)
        self.S_paths = sorted(make_dataset(self.dir_S, opt.max_dataset_size))
        assert len(self.S_paths) == 1, "SingleSkitDataset class should be used with one image in sketch S_paths {}".format(self.S_paths)
        self.S_img = self._load_sketch(self.S_paths[0])
        self.M_img, self.M_paths = None, [None]
        if self.opt.use_bg_mask is True:
            assert os.path.exists(self.dir_M), "Cannot find valid path for binary mask, %s" % self.dir_M
            self.M_paths = sorted(make_dataset(self.dir_M, opt.max_dataset_size))
            assert len(self.M_paths) == 1, "SingleSkitDataset class should be used with one image for mask"
            self.M_img = load_image_u8(self.M_paths[0], "L", self.device)
        if not os.path.exists(self.dir_I):
            print("Warning: missing I data opt dataroot {}, opt subdir_I {}".format(opt.dataroot, opt.subdir_I))
            assert "edit" in opt.dataroot, "I and T data are required for original sketches"
            self.I_paths, self.I_img, self.T_paths, self.T_size = [], None, [], 0
        else:
            assert os.path.exists(self.dir_I) and os.path.exists(self.dir_T), "datasets directories are invalid, \n dir_I {} \n dir_T {}".format(
                self.dir_I, self.dir_T)
            self.I_paths = sorted(make_dataset(self.dir_I, opt.max_dataset_size))
            assert len(self.I_paths) == 1, "SingleSkitDataset class should be used with one image in sketch and visual image, S_paths {}, I_paths {}".format(
                self.S_paths, self.I_paths)
            assert opt.image_nc == 3, "Visual image should have RGB 3 channels"
            self.I_img = load_image_u8(self.I_paths[0], "RGB", self.device)
            self.T_paths = make_touch_image_dataset(self.dir_T, opt.max_dataset_size)
            self.T_size = len(self.T_paths)
        if opt.subdir_valT is not None:
            self.val_T_paths = make_touch_image_dataset(self.dir_valT, opt.max_dataset_size)
            self.val_T_size = len(self.val_T_paths)
        else:
            self.val_T_paths, self.val_T_size = None, 0
        self.touch = TouchSet(self.T_paths, self.device, self.PATCH) if self.T_size > 0 else None
        self.val_touch = TouchSet(self.val_T_paths, self.device, self.PATCH) if self.val_T_size > 0 else None
        self.materials = [_Material(self.S_paths[0], self.S_img, self.I_img, self.M_img, self.M_paths[0], self.touch, self.val_touch)]

    # ---- the zoomed and ratio-resized sources of one (material, zoom level): computed once, shared by every index that uses them
    def _sources(self, mi, zi):
        key = (mi, zi)
        if key in self._src_cache:
            return self._src_cache[key]
        m = self.materials[mi]
        method = LANCZOS
        srcs = {"S": m.S_img, "I": m.I_img, "M": m.M_img}
        srcs.update(m.extra)
        if "zoom" in self.opt.preprocess:
            sf_h, sf_w = self.zoom_levels_A[zi]
            srcs = {k: (zoom_img(v, sf_h, sf_w, method) if v is not None else None) for k, v in srcs.items()}
        else:
            sf_h = sf_w = 1
        crop = self.opt.crop_size
        w, h = size_of(srcs["S"])
        ratio = crop_resize_ratio((w, h), crop, crop)

        def ratio_resize(v):      # crop_img resizes every image by the sketch's ratio, from the image's own size (dataset_util.py:186-194)
            vw, vh = size_of(v)
            return resize_u8(v, int(round(vh * ratio)), int(round(vw * ratio)), method)      # same size -> copy, as Image.resize
        out = {"img": {k: (ratio_resize(v) if v is not None else None) for k, v in srcs.items()}, "sf_h": sf_h, "sf_w": sf_w, "ratio": ratio}
        nbytes = sum(v.numel() for v in out["img"].values() if v is not None)
        if len(self._src_cache) == 0 or self._cache_used + nbytes <= self._cache_budget:
            self._src_cache[key] = out
            self._cache_used += nbytes
        return out

    def _final_u8(self, src, crop_pos_x, crop_pos_y):
        """The uint8 image after crop (+ the power-of-2 resize when the crop size is not a multiple of 256)."""
        crop = self.opt.crop_size
        h, w, c = src.shape
        out = torch.zeros((crop, crop, c), dtype=torch.uint8, device=self.device)       # PIL's crop pads with 0 outside the image
        ye, xe = min(h, crop_pos_y + crop), min(w, crop_pos_x + crop)
        out[:ye - crop_pos_y, :xe - crop_pos_x] = src[crop_pos_y:ye, crop_pos_x:xe]
        if (self.p2_w, self.p2_h) != (crop, crop):
            out = resize_u8(out, self.p2_h, self.p2_w, LANCZOS)
        return out

    def _image_tensor(self, src, normalize, crop_pos_x, crop_pos_y):
        crop = self.opt.crop_size
        if (self.p2_w, self.p2_h) == (crop, crop):
            return crop_to_tensor(src, crop_pos_x, crop_pos_y, crop, crop, normalize)
        u8 = self._final_u8(src, crop_pos_x, crop_pos_y)
        return crop_to_tensor(u8, 0, 0, self.p2_w, self.p2_h, normalize)

    def preprocess_data(self, timing=False, verbose=False, separate_val_set=False):
        """singleskit_dataset.py:194-432 / skit_dataset.py:211-500: draws every augmentation's crop and touch-patch selection (the
        reference's `random` call order), keeps the small tensors; the full-resolution tensors are produced in `__getitem__`."""
        if "padded" in self.opt.dataroot:
            self.padded_size = int(self.opt.dataroot.split("padded_")[1].split("/")[0].split("_")[0])
        crop = self.opt.crop_size
        self.p2_w, self.p2_h, self.resize_ratio_w, self.resize_ratio_h = make_power_2_size((crop, crop), 256)
        for index in range(len(self)):
            mi, zi = self._material_index(index), self._zoom_index(index)
            m = self.materials[mi]
            src = self._sources(mi, zi)
            W_, H_ = size_of(m.S_img)
            H, W = W_, H_            # the reference's `H, W = S_img.size[:2]` (PIL size is (width, height)): kept as is
            center_crop = "crop" not in self.opt.preprocess
            crop_pos_x, crop_pos_y = get_params(size_of(src["img"]["S"]), crop_size_h=crop, crop_size_w=crop, center_w=self.opt.center_w,
                                                center_h=self.opt.center_h, center_crop=center_crop)
            augmentation_params = {
                "H": H, "W": W,
                "scale_factor_h": src["sf_h"], "scale_factor_w": src["sf_w"],
                "crop_size_h": crop, "crop_size_w": crop,
                "resize_ratio": src["ratio"], "crop_pos_x": crop_pos_x, "crop_pos_y": crop_pos_y,
                "resize_ratio_w": self.resize_ratio_w, "resize_ratio_h": self.resize_ratio_h,
                "patch_crop_size": self.PATCH,
            }
            item = {"name": self._name(m), "S_paths": self.S_paths[0], "augmentation_params": augmentation_params}
            M3 = S3 = None
            if m.I_img is not None:
                if m.touch is not None or m.val_touch is not None:
                    M3 = self._final_u8(src["img"]["M"], crop_pos_x, crop_pos_y)
                    S3 = self._final_u8(src["img"]["S"], crop_pos_x, crop_pos_y)
                T_images, T_coords, full_T_coords, I_masks = [], [], [], []
                if m.touch is not None:
                    T_images, T_coords, full_T_coords, I_masks = self.find_validate_touch_patches_and_coords(
                        m.touch, augmentation_params, S3, M3, is_train=self.opt.is_train, is_val=False)
                val_T_images, val_T_coords, val_full_T_coords, val_I_masks = [], [], [], []
                if m.val_touch is not None:
                    val_T_images, val_T_coords, val_full_T_coords, val_I_masks = self.find_validate_touch_patches_and_coords(
                        m.val_touch, augmentation_params, S3, M3, is_train=self.opt.is_train, is_val=True)
                item.update({"I_masks": I_masks, "val_I_masks": val_I_masks, "T_images": T_images, "T_coords": T_coords,
                             "full_T_coords": full_T_coords, "val_T_images": val_T_images, "val_T_coords": val_T_coords,
                             "val_full_T_coords": val_full_T_coords})
            else:
                item["T_images"] = []
            if m.M_img is not None:
                item["M_paths"] = m.M_path
                if M3 is not None:
                    # extension (ignored by the reference model): the candidate table of get_patch_in_input's random fake patches for
                    # this item's mask, so that the B200 model's set_input does no device work for it
                    from .model_utils import offset_table_arrays
                    bits, rowcount = offset_table_arrays(M3[:, :, 0])
                    item["M_box_bits"], item["M_box_rowcount"] = torch.from_numpy(bits.copy()), torch.from_numpy(rowcount)
            self._index_key[index] = (mi, zi)
            self.data_dict[index] = item

    def find_validate_touch_patches_and_coords(self, touch, augmentation_params, S3, M3, is_train=False, is_val=False):
        """singleskit_dataset.py:434-660 + process_all_valid_patches (:662-1128) for one augmentation and one touch set."""
        a = augmentation_params
        rois1, rois3, valid_indexes = [], [], []
        for i in range(touch.P):
            ROI_x, ROI_y, ROI_h, ROI_w = touch.roi[i]
            if "padded" in self.opt.dataroot:
                ROI_x, ROI_y, ROI_h, ROI_w = global_padding_find_coords(ROI_x, ROI_y, ROI_h, ROI_w, padded_size=self.padded_size,
                                                                        org_h=self.opt.center_h, org_w=self.opt.center_w)
            x1, y1, h1, w1 = zoom_find_coords(ROI_x, ROI_y, ROI_h, ROI_w, scale_factor_h=a["scale_factor_h"], scale_factor_w=a["scale_factor_w"])
            ok, x2, y2, h2, w2 = crop_find_coords(x1, y1, h1, w1, a["crop_size_h"], a["crop_size_w"], a["resize_ratio"], a["crop_pos_x"], a["crop_pos_y"])
            x3, y3, h3, w3 = make_power_2_find_coords(x2, y2, h2, w2, a["resize_ratio_w"], a["resize_ratio_h"])
            if ok:
                valid_indexes.append(i)
                rois3.append([int(round(x3)), int(round(y3)), int(round(h3)), int(round(w3))])
        if valid_indexes != list(range(len(valid_indexes))):
            # the reference looks its valid-patch lists up by the RAW patch index (singleskit_dataset.py:742-743, 751): with a rejected
            # patch in front of a valid one it reads the wrong entry or runs off the list
            raise IndexError("a touch patch falls outside the augmented crop; the reference's valid-patch bookkeeping "
                             "(singleskit_dataset.py:742-751) only supports data whose patches are all inside every crop")
        nv = len(valid_indexes)
        if nv == 0:
            raise ValueError("no valid touch patch for this augmentation (the reference fails on the empty selection too)")
        # contact centres of every patch, in one batched pass (patches beyond the valid prefix do not exist, see above)
        roi_x = [r[0] for r in rois3] + [0] * (touch.P - nv)
        roi_y = [r[1] for r in rois3] + [0] * (touch.P - nv)
        counts, in_mask = touch.contact_centers(M3, roi_x, roi_y)
        # `np.sum(M3_arr[y:y+h, x:x+w]) == 0` looks at the ROI rectangle; the kernel looked at the patch's own h x w footprint at the
        # same origin — identical when the ROI keeps the patch size (no zoom); otherwise redo the test on the rectangle
        sel_patch, sel_rank, coords = [], [], []
        full_T_coords = []
        M3_2d = M3[:, :, 0]
        for i in range(nv):
            x3, y3, h3, w3 = rois3[i]
            if (h3, w3) != tuple(touch.hw[i]):
                in_rect = bool(M3_2d[max(0, y3):max(0, y3 + h3), max(0, x3):max(0, x3 + w3)].any().item())
            else:
                in_rect = bool(in_mask[i])
            if not in_rect:
                continue
            full_T_coords.append(rois3[i])
            n_c = int(counts[i])
            num_sample_bbox = min(n_c, self.opt.sample_bbox_per_patch)
            if is_train:
                selected = random.sample(range(n_c), num_sample_bbox)
            else:
                selected = np.arange(n_c // 2, n_c // 2 + num_sample_bbox)
            for k in selected:
                sel_patch.append(i); sel_rank.append(int(k))
        K = len(sel_patch)
        if K == 0:
            raise ValueError("no touch square survives the contact / object-mask test for this augmentation")
        cx, cy, all_T_images, all_I_masks = touch.squares(M3, sel_patch, sel_rank)
        half = touch.patch // 2
        for k in range(K):
            x3, y3, h3, w3 = rois3[sel_patch[k]]
            coords.append([x3, y3, h3, w3, a["patch_crop_size"], 1, int(cx[k]) - half, int(cy[k]) - half])
        all_T_coords = np.stack(coords, axis=0) if K > 1 else np.array(coords)

        calc_weight = bool(getattr(self.opt, "w_resampling", False))
        weights = None
        if calc_weight and is_train and not is_val:
            # offset = round(ROI + crop_pos), cutout = 32 (singleskit_dataset.py:1004-1012 with ratio 1, multiplier 1)
            var = laplacian_var(S3, [c[0] + c[6] for c in coords], [c[1] + c[7] for c in coords], a["patch_crop_size"], ref=255)
            weights = np.array([min(max(self.opt.resampling_w_min, v), self.opt.resampling_w_max) for v in var])

        total = K
        bs = min(self.opt.batch_size_G2, total) if getattr(self.opt, "batch_size_G2", 0) > 0 else total
        bs_val = min(self.opt.batch_size_G2_val, total) if getattr(self.opt, "batch_size_G2_val", 0) > 0 else total
        if is_train:
            if not is_val:
                if calc_weight:
                    selected_idxes = random.choices(range(total), weights=weights, k=bs)
                else:
                    selected_idxes = random.sample(range(total), bs)
            else:
                selected_idxes = random.sample(range(total), bs_val)
        else:
            print("test set, select all patches")
            selected_idxes = list(range(total))
        sel = torch.tensor(selected_idxes, dtype=torch.int64, device=self.device)
        return all_T_images[sel], all_T_coords[selected_idxes], full_T_coords, all_I_masks[sel]

    # ---- items
    def _images_for(self, index):
        if index in self._image_cache:
            return self._image_cache[index]
        a = self.data_dict[index]["augmentation_params"]
        img = self._sources(*self._index_key[index])["img"]
        out = {}
        for key, src in img.items():
            if src is not None:
                out[key] = self._image_tensor(src, not key.endswith("M"), a["crop_pos_x"], a["crop_pos_y"])     # masks: ToTensor only
        nbytes = sum(t.numel() * 4 for t in out.values())
        if self._cache_used + nbytes <= self._cache_budget:
            self._image_cache[index] = out
            self._cache_used += nbytes
        return out

    def __getitem__(self, index):
        assert index in self.data_dict.keys(), "Cannot find index %d in dataset" % (index)
        item = dict(self.data_dict[index])
        item.update(self._images_for(index))
        if self.host_items:
            if index not in self._host_cache:      # one device -> host copy per item, then the same host tensors every epoch
                self._host_cache[index] = {k: v.cpu() for k, v in item.items() if torch.is_tensor(v)}
            item.update(self._host_cache[index])
        return item

    def __len__(self):
        return self.data_len


class SkitDataset(SingleSkitDataset):
    """data/skit_dataset.py:25 — the multi-material dataset of the skitG model: `opt.material_list` objects, item `index` belongs to
    material `index % len(material_list)` (:240) and uses its own zoom level `zoom_levels_A[index]` (:278).  Each material's directory
    is `<datasets_dir>/singleskit_<material>_padded_<padded_size>_x<T_resolution_multiplier>/` (:141; `datasets_dir` = `./datasets`
    unless `opt.datasets_dir` says otherwise)."""

    def _material_index(self, index):
        return index % len(self.opt.material_list)

    def _zoom_index(self, index):
        return index

    def _name(self, m):
        return os.path.splitext(ntpath.basename(m.S_path))[0]

    def _load_materials(self):
        """skit_dataset.py:86-196."""
        opt = self.opt
        base = getattr(opt, "datasets_dir", "./datasets")
        if not getattr(opt, "load_contact_mask", True):
            raise NotImplementedError("load_contact_mask=False (PIL random square crops of the touch maps, singleskit_dataset.py:862-905) "
                                      "is not built on the B200 path")
        if hasattr(opt, "material_list"):
            print("material_list is {}".format(opt.material_list))
        self.S_paths, self.I_paths, self.M_paths, self.T_paths, self.T_sizes, self.val_T_paths, self.val_T_sizes = [], [], [], [], [], [], []
        external = bool(getattr(opt, "use_external_test_input", False))
        if external:
            self.style_I_paths, self.style_M_paths = [], []
            print("Use external test input")
            assert hasattr(opt, "test_sketch_material"), "test_sketch_material is not defined"
            sketch_root = os.path.join(base, f"singleskit_{opt.test_sketch_material}_padded_{opt.padded_size}_x{opt.T_resolution_multiplier}_edit0/")
            self.S_paths.extend(sorted(make_dataset(os.path.join(sketch_root, opt.subdir_S), opt.max_dataset_size)))
            self.M_paths.extend(sorted(make_dataset(os.path.join(sketch_root, opt.subdir_M), opt.max_dataset_size)))
            style_root = os.path.join(base, f"singleskit_{opt.test_style_material}_padded_{opt.padded_size}_x{opt.T_resolution_multiplier}_edit0/")
            self.style_I_paths.extend(sorted(make_dataset(os.path.join(style_root, opt.subdir_I), opt.max_dataset_size)))
            self.style_M_paths.extend(sorted(make_dataset(os.path.join(style_root, opt.subdir_M), opt.max_dataset_size)))
        else:
            print("Iterate over material_list")
            for material in opt.material_list:
                dataroot = os.path.join(base, f"singleskit_{material}_padded_{opt.padded_size}_x{opt.T_resolution_multiplier}/")
                dir_S, dir_I = os.path.join(dataroot, opt.subdir_S), os.path.join(dataroot, opt.subdir_I)
                dir_T, dir_M = os.path.join(dataroot, opt.subdir_T), os.path.join(dataroot, opt.subdir_M)
                dir_valT = os.path.join(dataroot, opt.subdir_valT) if opt.subdir_valT is not None else None
                assert os.path.exists(dir_S) and os.path.exists(dir_I) and os.path.exists(dir_T) and os.path.exists(dir_M), \
                    "datasets directories are invalid, \n dir_S {} \n dir_I {} \n dir_T {} \n dir_M {}".format(dir_S, dir_I, dir_T, dir_M)
                self.S_paths.extend(sorted(make_dataset(dir_S, opt.max_dataset_size)))
                self.I_paths.extend(sorted(make_dataset(dir_I, opt.max_dataset_size)))
                self.M_paths.extend(sorted(make_dataset(dir_M, opt.max_dataset_size)))
                T_paths = make_touch_image_dataset(dir_T, opt.max_dataset_size)
                self.T_paths.append(T_paths)
                self.T_sizes.append(len(T_paths))
                val_T_paths = make_touch_image_dataset(dir_valT, opt.max_dataset_size) if dir_valT is not None else []
                self.val_T_paths.append(val_T_paths)
                self.val_T_sizes.append(len(val_T_paths))
        assert opt.image_nc == 3, "Visual image should have RGB 3 channels"
        self.materials = []
        for k, S_path in enumerate(self.S_paths):
            I_img = load_image_u8(self.I_paths[k], "RGB", self.device) if len(self.I_paths) > 0 else None
            M_img = load_image_u8(self.M_paths[k], "L", self.device) if opt.use_bg_mask is True else None
            extra = None
            if external:
                extra = {"style_I": load_image_u8(self.style_I_paths[k], "RGB", self.device),
                         "style_M": load_image_u8(self.style_M_paths[k], "L", self.device)}
            touch = TouchSet(self.T_paths[k], self.device, self.PATCH) if I_img is not None and self.T_sizes[k] > 0 else None
            val_touch = TouchSet(self.val_T_paths[k], self.device, self.PATCH) if I_img is not None and self.val_T_sizes[k] > 0 else None
            self.materials.append(_Material(S_path, self._load_sketch(S_path), I_img, M_img, self.M_paths[k] if opt.use_bg_mask else None,
                                            touch, val_touch, extra))
