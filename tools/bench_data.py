"""Data-pipeline benchmark at the reference's scale: one padded 1800 x 1800 sketch / image / mask, crop 1536,
`data_len` augmentations, 60 + 20 touch patches of ~180 x 230 pixels with ~2 000 contact-centre pixels each — the work the reference
does at start-up in `SingleSkitDataset.preprocess_data` ("20-30 min", README.md:129).

    python tools/bench_data.py [--data-len 200] [--cpu-items 2]          # device dataset vs the numpy oracle port (bounded sample)
    python tools/bench_data.py --impl reference --cpu-items 2             # the unmodified reference class (build container only)

Prints one JSON line: seconds per augmentation for each arm (same unit, nothing extrapolated), device build time, item fetch time.
"""
import argparse
import json
import os
import random
import sys
import time

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import make_data_golden as MG  # noqa: E402


def big_dataset(root):
    if not os.path.exists(os.path.join(root, "trainS", "syn.png")):
        MG.synth_dataset(root, seed=1, W=1800, H=1800, n_train=60, n_val=20, patch_h=(160, 200), patch_w=(200, 260), patch_x=(0, 1000),
                         patch_y=(0, 700), strokes=2500)
    return root


def options(root, data_len):
    return MG.dataset_options(root, crop_size=1536, center_w=1280, center_h=960, data_len=data_len, batch_size_G2=64, batch_size_G2_val=128,
                              sample_bbox_per_patch=2)


def _time(fn, reps=20):
    import torch
    for _ in range(3):
        fn()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps * 1e-3


def kernel_rooflines(ds):
    """The data kernels against the HBM roofline (all are byte movers): algorithmic bytes = source read once + result written once."""
    import torch
    from vts_b200 import data_pipeline as DP
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
        src = "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        peak, src = 6500.0, "fallback 6.5 TB/s"
    I = ds.I_img                                            # 1800 x 1800 x 3 uint8
    h, w, c = I.shape
    oh, ow = int(round(h * 0.8)), int(round(w * 0.9))
    res = {}
    t = _time(lambda: DP.resize_u8(I, oh, ow, DP.LANCZOS))
    byt = h * w * c + h * ow * c * 2 + oh * ow * c          # source, the horizontal pass's intermediate (written + read), result
    res["resize_u8 LANCZOS 1800x1800x3 -> %dx%d" % (oh, ow)] = dict(us=t * 1e6, gbs=byt / t / 1e9, frac=byt / t / 1e9 / peak)
    crop = ds.opt.crop_size
    t = _time(lambda: DP.crop_to_tensor(I, 100, 120, crop, crop, True))
    byt = crop * crop * c * (1 + 4)
    res["u8_crop_to_tensor %dx%dx3" % (crop, crop)] = dict(us=t * 1e6, gbs=byt / t / 1e9, frac=byt / t / 1e9 / peak)
    M3 = ds._final_u8(ds._sources(0, 0)["img"]["M"], 100, 120)
    ts = ds.touch
    rx = [int(r[0]) % 1000 for r in ts.roi]; ry = [int(r[1]) % 700 for r in ts.roi]
    t = _time(lambda: ts.contact_centers(M3, rx, ry))
    byt = ts.total * (8 + 1 + 3 + 3 + 4)                    # contact mask fp64 + centre mask + three byte maps written and read back + centre list
    res["contact_centers %d patches, %d pixels (incl. the count read-back)" % (ts.P, ts.total)] = dict(us=t * 1e6, gbs=byt / t / 1e9, frac=byt / t / 1e9 / peak)
    return {"peak_gbs": peak, "peak_source": src, "kernels": res}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--data-len", type=int, default=200)
    ap.add_argument("--cpu-items", type=int, default=2)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--root", default="/tmp/vts_bench_data/singleskit_syn_padded_1800_x1")
    a = ap.parse_args()
    root = big_dataset(a.root)
    out = {"workload": "SingleSkitDataset.preprocess_data: 1800x1800 padded sources, crop 1536, 60 train + 20 val touch patches (~180x230, ~2000 centre "
                       "pixels each), sample_bbox_per_patch 2, batch_size_G2 64 / 128, w_resampling on"}
    if a.impl == "reference":
        opt = options(root, a.cpu_items)
        t0 = time.time()
        MG.run_reference(root, opt, seed=7)
        out.update(impl="reference", items=a.cpu_items, s_per_augmentation=(time.time() - t0) / a.cpu_items)
        print(json.dumps(out))
        return
    import torch
    import vts_b200
    opt = options(root, 2)
    random.seed(7); np.random.seed(7)
    vts_b200.SingleSkitDataset(opt)                       # warm-up: library load, allocator, file cache
    opt = options(root, a.data_len)
    random.seed(7); np.random.seed(7)
    torch.cuda.synchronize(); t0 = time.time()
    ds = vts_b200.SingleSkitDataset(opt, cache_bytes=0)
    torch.cuda.synchronize(); t_build = time.time() - t0
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for i in range(a.data_len):
        item = ds[i]
    e1.record(); torch.cuda.synchronize()
    out.update(impl="b200", data_len=a.data_len, build_s=t_build, s_per_augmentation=t_build / a.data_len,
               item_fetch_ms=e0.elapsed_time(e1) / a.data_len, item_bytes=sum(v.numel() * v.element_size() for v in item.values() if torch.is_tensor(v)))
    out["kernels"] = kernel_rooflines(ds)
    if a.cpu_items > 0:
        from oracle import data_oracle as DO
        opt = options(root, a.cpu_items)
        random.seed(7); np.random.seed(7)
        t0 = time.time()
        items = DO.dataset_items(opt)
        t_cpu = (time.time() - t0) / a.cpu_items
        out["cpu_baseline"] = {"kind": "port", "what": "oracle/data_oracle.py dataset_items (numpy restatement of the reference class)", "cores": 1,
                               "items": a.cpu_items, "s_per_augmentation": t_cpu}
        # parity at the benchmark's own size: same seed -> same first items
        from oracle.data_oracle import to_tensor_norm
        checks = {"T_images": lambda i: np.array_equal(ds[i]["T_images"].cpu().numpy(), items[i]["T_images"]),
                  "T_coords": lambda i: np.array_equal(np.asarray(ds[i]["T_coords"]), items[i]["T_coords"]),
                  "I_masks": lambda i: np.array_equal(ds[i]["I_masks"].cpu().numpy(), items[i]["I_masks"]),
                  "val_T_images": lambda i: np.array_equal(ds[i]["val_T_images"].cpu().numpy(), items[i]["val_T_images"]),
                  "S": lambda i: np.array_equal(ds[i]["S"].cpu().numpy(), to_tensor_norm(items[i]["S_u8"])),
                  "I": lambda i: np.array_equal(ds[i]["I"].cpu().numpy(), to_tensor_norm(items[i]["I_u8"])),
                  "M": lambda i: np.array_equal(ds[i]["M"].cpu().numpy(), to_tensor_norm(items[i]["M_u8"], normalize=False))}
        bad = [(k, i) for k, f in checks.items() for i in range(a.cpu_items) if not f(i)]
        same = not bad
        if bad:
            out["mismatches"] = bad
        out["matches_oracle"] = bool(same)
        out["speedup_vs_cpu_port"] = t_cpu / (t_build / a.data_len)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
