"""bench.py — skitG/sinskitG train-step throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--no-nce] [--impl b200|reference|eager]

Workload (config.workload): BASELINE.json configs[2], the configuration `north_star` quotes throughput on — one
single-material train step (G forward, D1 step, D2 step, G step with GAN + L1 + patch-L1 + PatchNCE, three Adam updates) at
768x768, NT = 64 touch patches + NF = 32 random fake patches, PatchNCE on (5 feature layers, 256 patches, T = 0.07), LPIPS /
vision-aided off, on the tensor-core architecture (resnet_9blocks ngf 64, multiscale PatchGAN ndf 64), batch 1 per rank (the
reference forces batch_size = 1).  Synthetic seeded inputs, random-init weights.  `--size 512 --no-nce` is configs[1]; it is
also measured (shorter) into `extra.config1_512` of the default line.

One JSON line on stdout (rank 0).  `value` = images/s with inputs resident in HBM; `e2e` = the same step through the public
model API with host inputs: set_input (pinned H2D) + optimize_parameters + get_current_losses (D2H) inside the timed region.
N > 1: one process per GPU (torchrun), one sample per rank per step, flat-bucket gradient all-reduce over NCCL — weak scaling.

`--impl reference`: the reference's own CPU implementation of the SAME workload (same size, NT, NF, PatchNCE) on the host cores,
never scaled: the unmodified reference through oracle/ref_step.py when its tree is mounted and the workload is one it
implements (PatchNCE is dead code in the reference, so configs[2] always runs the oracle port), else the oracle port
(oracle/skit_oracle.py, pinned to the real reference by tests/golden).  This arm never imports the product package.
`--impl eager`: the same oracle step with every tensor on cuda:0 — eager PyTorch + cuDNN on the same B200 (TF32 convolutions on,
torch's default, and off), the bar SURVEY.md section 2.3 names.  A short eager run is also folded into the default line
(`eager_b200`) at N = 1.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402
import torch  # noqa: E402

METRIC = "skitG train-step images/sec"
UNIT = "images/s"
NT, NF = 64, 32


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=30)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=768)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference", "eager"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra", action="store_true", help="skip the extra measurements folded into the default line (configs[1], "
                                                            "arch A, eager PyTorch on the same GPU, 1024x1024 inference)")
    ap.add_argument("--mode", default="train", choices=["train", "infer"], help="infer: generator forward only (BASELINE.json configs[4])")
    ap.add_argument("--batch", type=int, default=1, help="images per rank per step (infer mode)")
    ap.add_argument("--nce", dest="nce", action="store_true", default=True, help="PatchNCE on (default; BASELINE.json configs[2])")
    ap.add_argument("--no-nce", dest="nce", action="store_false", help="PatchNCE off (configs[1] wiring: use with --size 512)")
    ap.add_argument("--materials", type=int, default=1, help="distinct synthetic materials cycled round-robin over ranks and steps "
                                                             "(BASELINE.json configs[3]: 20)")
    ap.add_argument("--lpips", action="store_true", help="LPIPS-VGG16 terms on with the reference's default weights (lambda_G1_lpips 1, "
                                                         "lambda_G2_lpips 10; random VGG weights: no checkpoint offline)")
    return ap.parse_args()


def config_index(a):
    if a.materials > 1:
        return 3
    return 2 if a.nce else 1


def workload_config(a):
    return {"workload": "skitG/sinskitG train step, %s, %dx%d, resnet_9blocks ngf64 + multiscale PatchGAN ndf64, "
                        "NT=64 NF=32, GAN+L1+patch-L1, PatchNCE %s, LPIPS %s, VAL off (BASELINE.json configs[%d]%s)"
                        % ("single material" if a.materials == 1 else "%d materials round-robin over ranks" % a.materials,
                           a.size, a.size, "on (5 layers, 256 patches, T=0.07)" if a.nce else "off",
                           "on (VGG16, full image + touch patches, random weights)" if a.lpips else "off", config_index(a),
                           " + the reference's default LPIPS terms" if a.lpips else ""),
            "size": a.size, "images_per_rank_per_step": 1, "NT": NT, "NF": NF, "netG": "resnet_9blocks", "ngf": 64,
            "netD": "multiscale", "ndf": 64, "patchnce": bool(a.nce), "materials": a.materials,
            "parallelism": "dp%d (flat grad bucket all-reduce, NCCL)" % a.gpus,
            "l2": "per-step working set (saved activations + operands) is several GB >> 126 MB L2; no flush needed"}


# ----------------------------------------------------------------------------------------- clocks
class ClockSampler:
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.stop = index, [], threading.Event()
        self.t = threading.Thread(target=self.run, daemon=True)

    def run(self):
        while not self.stop.is_set():
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(",")])
            except Exception:
                pass
            self.stop.wait(0.05)

    def __enter__(self):
        self.t.start()
        return self

    def __exit__(self, *a):
        self.stop.set()
        self.t.join(timeout=6)

    def summary(self):
        if not self.rows:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["unsampled"]}
        sm = sorted(int(r[0]) for r in self.rows if r[0].isdigit())
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 2 + i and r[2 + i].lower().startswith("active") for r in self.rows)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": int(self.rows[0][1]) if self.rows[0][1].isdigit() else None,
                "reasons": reasons, "samples": len(self.rows)}


# ----------------------------------------------------------------------------------------- the reference's own code path (CPU / eager GPU)
def oracle_stepper(size, nce, lpips, device="cpu", seed=0, arch="B"):
    """One train step of the oracle (oracle/skit_oracle.py: the reference's own ATen code path restated call for call and pinned
    to the real reference by tests/golden) at the FULL workload: size x size, NT 64, NF 32.  Weights come from the oracle's own
    init tables — the product package is not imported.  -> step(i) callable."""
    from oracle import skit_oracle as O
    if arch == "A":      # the reference's default architecture: unet256_custom ngf 10, multiscale PatchGAN ndf 8
        sds = [O.init_unet_custom(9, 10, 8, 4, seed=seed), O.init_multiscale_d(4, 8, 3, 3, seed=seed + 1), O.init_multiscale_d(7, 8, 3, 3, seed=seed + 2)]
    else:
        sds = [O.init_resnet_g(9, 5, 64, 9, seed=seed), O.init_multiscale_d(4, 64, 3, 3, seed=seed + 1), O.init_multiscale_d(7, 64, 3, 3, seed=seed + 2)]
    sds = [{k: v.to(device) for k, v in sd.items()} for sd in sds]
    cfg = O.StepConfig(netG="unet256_custom" if arch == "A" else "resnet_9blocks", batch_size_G2=NT, add_fake_T_sample_size=NF, lambda_NCE=1.0 if nce else 0.0,
                       lambda_G1_lpips=1.0 if lpips else 0.0, lambda_G2_lpips=10.0 if lpips else 0.0, foreach_adam=device != "cpu")
    sdL = {k: v.to(device) for k, v in O.lpips_random_state(0).items()} if lpips else None
    batch = O.step_inputs_from_batch(O.synthetic_batch(size, NT=NT, seed=seed), device=None if device == "cpu" else device)
    rs = np.random.RandomState(seed)
    feat_hw = {0: (size + 6, size + 6), 4: (size, size), 8: (size // 2, size // 2), 12: (size // 4, size // 4), 16: (size // 4, size // 4)}
    state = {}

    def step(i):
        rand = dict(real_b=[rs.rand()], real_s=[rs.rand()], fake_b=[rs.rand()], fake_s=[rs.rand()],
                    fake_ox=rs.randint(0, size - 32, NF).astype(np.int32), fake_oy=rs.randint(0, size - 32, NF).astype(np.int32))
        if nce:     # PatchSampleF's own draw (networks.py:703-705): a full permutation per layer per step
            rand["nce_ids"] = [np.random.permutation(h * w)[:min(cfg.num_patches, h * w)] for h, w in (feat_hw[l] for l in cfg.nce_layers)]
        return O.train_step(cfg, sds[0], sds[1], sds[2], state, batch, rand, step=i + 1, sdL=sdL)["losses"]

    return step


def time_cpu_path(a, steps, warmup):
    """-> (seconds per step, kind, sample description, cores)."""
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    kind, what, step = "port", "the oracle port of the reference's train step (oracle/skit_oracle.py)", None
    if not a.nce and not a.lpips:
        try:
            from oracle import ref_step
            if ref_step.reference_available():
                run, what = ref_step.reference_step_runner(a.size, NT, NF)
                step, kind = (lambda i: run()), "reference"
        except Exception as e:      # the reference tree is read-only test infrastructure: fall back to the port, say why
            what += " [reference unavailable: %s]" % str(e)[:80]
    if step is None:
        step = oracle_stepper(a.size, a.nce, a.lpips)
    times = []
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        step(i)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    t = float(np.median(times))
    sample = "%s at the full workload %dx%d NT=%d NF=%d PatchNCE %s: %d warm-up + %d timed steps, median %.2f s/step, %d threads" % (
        what, a.size, a.size, NT, NF, "on" if a.nce else "off", warmup, steps, t, cores)
    return t, kind, sample, cores


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    t, kind, sample, cores = time_cpu_path(a, max(1, min(a.steps, 3)), max(1, min(a.warmup, 1)))
    value = 1.0 / t
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1000.0 * t, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(a),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


def time_eager_gpu(size, nce, lpips, steps=3, warmup=1, arch="B"):
    """The reference's code path as eager PyTorch + cuDNN on this GPU: the oracle's train step with every tensor on cuda:0,
    autograd backward, multi-tensor Adam; CUDA-event timed.  TF32 convolutions on (torch's default) and off (the fp32 the parity
    gate is defined against)."""
    out = {}
    for tag, tf32 in (("tf32", True), ("fp32", False)):
        old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark)
        torch.backends.cudnn.allow_tf32 = tf32
        torch.backends.cuda.matmul.allow_tf32 = False
        torch.backends.cudnn.benchmark = True       # base_model.py:38
        try:
            step = oracle_stepper(size, nce, lpips, device="cuda", arch=arch)
            for i in range(warmup):
                step(i)
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(steps):
                step(warmup + i)
            e1.record()
            torch.cuda.synchronize()
            ms = e0.elapsed_time(e1) / steps
            out[tag] = {"ms_per_step": ms, "images_per_s": 1e3 / ms}
        finally:
            torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32, torch.backends.cudnn.benchmark = old
        torch.cuda.empty_cache()
    out["what"] = ("oracle train step (the reference's ATen call sequence + autograd + multi-tensor Adam) with all tensors on cuda:0, %dx%d NT=%d NF=%d "
                   "PatchNCE %s, cudnn.benchmark on, %d warm-up + %d timed steps, CUDA events" % (size, size, NT, NF, "on" if nce else "off", warmup, steps))
    return out


def run_eager(a):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    torch.cuda.set_device(0)
    with ClockSampler(0) as clk:
        r = time_eager_gpu(a.size, a.nce, a.lpips, steps=max(1, min(a.steps, 10)), warmup=max(1, min(a.warmup, 3)))
    v = r["tf32"]["images_per_s"]
    line = {"impl": "eager", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": 1, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": r["tf32"]["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "tf32 convolutions (torch default), fp32 elsewhere", "data": "synthetic", "config": workload_config(a),
            "eager_b200": r, "clocks": clk.summary()}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- the B200 arm
def _ncu_field(size, field):
    """A per-launch figure of the dominant kernel from the committed `ncu --set full` summary under profiles/ (newest round first)."""
    for rnd in ("r02f", "r02", "r01"):
        try:
            prof = json.load(open(os.path.join(ROOT, "profiles", "%s_conv_tc_halo_fwd%d.json" % (rnd, size))))
            nested = {"tensor_pipe_pct_elapsed": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed",
                      "tensor_pipe_pct_active": "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"}.get(field)
            val = prof.get(field)
            if val is None and nested in prof:
                val = float(str(prof[nested]["value"]).replace(",", ""))
            if val is not None:
                return val, "%s_conv_tc_halo_fwd%d.json" % (rnd, size)
        except Exception:
            pass
    return None, None


def measure_dominant_kernel(size, peaks):
    """The dominant kernel of the step is the tcgen05 ResnetBlock conv (256->256, 3x3, (S/4)^2 pixels; the same kernel
    runs its input gradient).  Timed alone with CUDA events around a replayed CUDA graph of launches (device time, no
    host launch cost) that rotate over operand/output buffer pairs (> 126 MB in total, so no launch finds its input
    or output resident in L2 from the previous one); algorithmic FLOP/s against the measured bf16 peak."""
    import math
    from vts_b200 import ops
    s = size // 4
    nbuf = max(2, int(math.ceil(160e6 / (s * s * 256 * 8.3))))
    w = torch.randn(256, 256, 3, 3, device="cuda") / math.sqrt(2304)
    pk = ops.PackedWeights(w, 0, want_f32=False, want_bf16=True)
    opsx, ys = [], []
    for _ in range(nbuf):
        x = torch.randn(1, s, s, 256, device="cuda")
        opsx.append(ops.norm_act_pad(x, pad=1, pad_mode=ops.PAD_REFLECT, fmt=ops.FMT_BF16X2)[1])
        ys.append(torch.empty(1, s, s, 256, device="cuda"))
    stats = torch.zeros(1, 256, 2, dtype=torch.float64, device="cuda")

    def launch(i):
        ops.L.call("skit_conv2d_fwd", opsx[i % nbuf].ref(), pk.ref(), 1, 0, s, s, None, ops._p(ys[i % nbuf]), ops._p(stats),
                   ops.NORM_INSTANCE, ops.IMPL_TC, ops.L.stream())

    for i in range(4):
        launch(i)
    torch.cuda.synchronize()
    iters = 3 * nbuf
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for i in range(iters):
            launch(i)
    g.replay()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / iters
    flops = 2.0 * 9 * 256 * 256 * s * s
    achieved = flops / (ms * 1e-3) / 1e12
    peak = float(peaks.get("bf16_tflops", 1590.0))
    traffic, src = _ncu_field(size, "dram_bytes_per_launch")
    pipe_el, _ = _ncu_field(size, "tensor_pipe_pct_elapsed")
    pipe_ac, _ = _ncu_field(size, "tensor_pipe_pct_active")
    return {"bound": "tensor", "kernel": "conv_tc_halo_kernel<256> (ResnetBlock conv3x3 256->256 @%dx%d)" % (s, s),
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "executed_frac": 3 * achieved / peak,
            "tensor_pipe_pct_elapsed": pipe_el, "tensor_pipe_pct_active": pipe_ac, "ncu_source": src,
            "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops, burst)" if "bf16_tflops" in peaks else "fallback 1.59 PFLOP/s",
            "traffic": traffic, "ms_per_launch": ms, "timing": "CUDA events around a replayed graph of %d launches over %d buffer pairs (> L2)" % (iters, nbuf),
            "note": "achieved / frac count ALGORITHMIC flops; the kernel executes 3 bf16 MMAs per product (hi/lo split, the fp32-parity "
                    "requirement: DESIGN.md 4.1), so frac is capped at 1/3 and executed_frac = 3 x frac is the tensor-pipe rate actually "
                    "sustained; tensor_pipe_pct_* are ncu's sm__pipe_tensor_cycles_active from the committed --set full capture (the "
                    "north_star's target metric: >= 60 % of elapsed)"}


def time_train_config(size, nce, steps, local_rank, seed, netG="resnet_9blocks", ngf=64, ndf=64):
    """A short resident-input measurement of another configuration (the `extra` entries)."""
    import vts_b200
    from oracle import skit_oracle as O
    opt = vts_b200.default_options(netG=netG, ngf=ngf, ndf=ndf, gpu_ids=[local_rank], lambda_NCE=1.0 if nce else 0.0)
    torch.manual_seed(0)
    m = vts_b200.SinSKITGModel(opt)
    m.set_input(O.synthetic_batch(size, NT=NT, seed=seed))
    for _ in range(4):
        m.optimize_parameters(1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        m.optimize_parameters(1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del m
    torch.cuda.empty_cache()
    return {"size": size, "patchnce": nce, "netG": netG, "ngf": ngf, "ndf": ndf, "steps": steps, "ms_per_step": ms, "images_per_s_per_gpu": 1e3 / ms}


def time_infer(size, B, steps, local_rank):
    import vts_b200
    opt = vts_b200.default_options(gpu_ids=[local_rank], isTrain=False)
    torch.manual_seed(0)
    m = vts_b200.SinSKITGModel(opt)
    g = torch.Generator().manual_seed(1)
    m.set_input({"S": torch.rand(B, 1, size, size, generator=g) * 2 - 1, "M": torch.ones(B, 1, size, size)}, phase="test")
    for _ in range(3):
        m.test()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(steps):
        m.test()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / steps
    del m
    torch.cuda.empty_cache()
    return {"size": size, "batch": B, "ms_per_step": ms, "images_per_s_per_gpu": 1e3 * B / ms}


def run_b200(a):
    import vts_b200
    from vts_b200 import _lib
    from vts_b200.dist import DistContext
    from oracle import skit_oracle as O  # synthetic batch factory only (seeded inputs of SURVEY.md §8d)
    ctx = DistContext()
    torch.cuda.set_device(ctx.local_rank)
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    opt = vts_b200.default_options(gpu_ids=[ctx.local_rank], lambda_NCE=1.0 if a.nce else 0.0,
                                   lambda_G1_lpips=1.0 if a.lpips else 0.0, lambda_G2_lpips=10.0 if a.lpips else 0.0,
                                   allow_random_lpips=a.lpips)
    torch.manual_seed(0)
    model = vts_b200.SinSKITGModel(opt, dist_ctx=ctx if ctx.world_size > 1 else None)
    ctx.broadcast_params([model.netG, model.netD, model.netD2])
    # a different (material, augmentation) sample per rank; --materials M: rank r cycles through materials r, r+W, ... (round-robin,
    # material_index = index % len(material_list), data/skit_dataset.py:240)
    mats = ctx.sample_indices(a.materials) if a.materials > 1 else [ctx.rank]
    if not mats:
        mats = [ctx.rank % a.materials]
    batches = []
    for mi in mats:
        b = O.synthetic_batch(a.size, NT=NT, seed=mi)
        for k in ("S", "I", "M", "T_images", "I_masks"):
            b[k] = b[k].pin_memory()
        batches.append(b)
    it = {"i": 0}

    def step_resident():
        if len(batches) > 1:       # several materials per rank: the inputs change every step (the H2D copy is part of the step)
            model.set_input(batches[it["i"] % len(batches)])
            it["i"] += 1
        model.optimize_parameters(1)

    def step_e2e():
        model.set_input(batches[it["i"] % len(batches)])
        it["i"] += 1
        model.optimize_parameters(1)
        return model.get_current_losses()

    model.set_input(batches[0])
    for _ in range(max(a.warmup, 3)):
        step_resident()
    # ---- timed region 1: inputs resident in HBM
    ctx.barrier()
    torch.cuda.synchronize()
    l0 = _lib.launches
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    with ClockSampler(ctx.local_rank) as clk:
        e0.record()
        for _ in range(a.steps):
            step_resident()
        e1.record()
        torch.cuda.synchronize()
    ctx.barrier()
    launches = _lib.launches - l0
    t_res = ctx.max_over_ranks(e0.elapsed_time(e1) / 1e3)
    # ---- timed region 2: end to end through the public API, host inputs
    step_e2e()
    ctx.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    e0.record()
    for _ in range(a.steps):
        losses = step_e2e()
    e1.record()
    torch.cuda.synchronize()
    t_wall = time.perf_counter() - t0
    ctx.barrier()
    t_e2e = ctx.max_over_ranks(max(e0.elapsed_time(e1) / 1e3, t_wall))
    d2h = 4 * (8 + 3 * NT + NF) + (4 * 5 if a.nce else 0)
    n = ctx.world_size
    value = n * a.steps / t_res
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": n, "steps": a.steps, "warmup": max(a.warmup, 3),
            "ms_per_step": 1e3 * t_res / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3 split operands, fp32 accumulate (fp32-parity tensor-core path); fp32 elsewhere",
            "data": "synthetic", "config": workload_config(a),
            "e2e": {"value": n * a.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(model.h2d_bytes), "d2h_bytes_per_step": d2h},
            "gpu_launches": int(launches), "clocks": clk.summary(),
            "losses_last_step": {k: round(v, 5) for k, v in losses.items()}}
    if ctx.rank == 0:
        # free the timed model's activations before the side measurements
        model._graph = None if n == 1 else model._graph
        line["roofline"] = measure_dominant_kernel(a.size, peaks)
        if n == 1 and not a.no_extra:
            extra = {}
            try:
                if a.nce or a.size != 512:
                    extra["config1_512"] = dict(time_train_config(512, False, 20, ctx.local_rank, ctx.rank), config="BASELINE.json configs[1]")
                extra["arch_A_default"] = dict(time_train_config(a.size, False, 20, ctx.local_rank, ctx.rank, netG="unet256_custom", ngf=10, ndf=8),
                                               config="the reference's default architecture (unet256_custom ngf 10, ndf 8), PatchNCE off")
                extra["infer_1024"] = [time_infer(1024, B, 6, ctx.local_rank) for B in (1, 8)]
                line["eager_b200"] = time_eager_gpu(a.size, a.nce, a.lpips, steps=3, warmup=1)
                extra["arch_A_default"]["eager_b200"] = time_eager_gpu(a.size, False, False, steps=5, warmup=2, arch="A")
            except Exception as e:       # a side measurement must never cost the headline line
                extra["error"] = repr(e)[:300]
            line["extra"] = extra
        if n == 1 and not a.no_cpu_baseline:
            t, kind, sample, cores = time_cpu_path(a, 1, 1)
            line["cpu_baseline"] = {"value": 1.0 / t, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample}
        print(json.dumps(line), flush=True)
    if ctx.world_size > 1:
        # Tearing NCCL down while a captured CUDA graph still holds its collectives can block forever in
        # destroy_process_group: drop the graph, drain the device, meet at a barrier, then leave without the
        # communicator teardown (the process is exiting anyway).
        model._graph = None
        torch.cuda.synchronize()
        ctx.barrier()
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    ctx.shutdown()


def run_infer(a):
    """BASELINE.json configs[4]: sketch -> (RGB, normal, height-gradient) generator forward throughput, batch B per rank,
    replicas only (no collective).  `value`: inputs resident; `e2e`: set_input (pinned H2D of S, M) + test() + D2H of fake_N."""
    import vts_b200
    from vts_b200.dist import DistContext
    ctx = DistContext()
    torch.cuda.set_device(ctx.local_rank)
    size = a.size
    opt = vts_b200.default_options(gpu_ids=[ctx.local_rank], isTrain=False)
    torch.manual_seed(0)
    model = vts_b200.SinSKITGModel(opt)
    g = torch.Generator().manual_seed(ctx.rank)
    B = a.batch
    batch = {"S": (torch.rand(B, 1, size, size, generator=g) * 2 - 1).pin_memory(), "M": torch.ones(B, 1, size, size).pin_memory()}
    out_host = torch.empty(B, 3, size, size).pin_memory()
    model.set_input(batch, phase="test")
    for _ in range(max(a.warmup, 3)):
        model.test()
    ctx.barrier()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = vts_b200._lib.launches
    with ClockSampler(ctx.local_rank) as clk:
        e0.record()
        for _ in range(a.steps):
            model.test()
        e1.record()
        torch.cuda.synchronize()
    launches = vts_b200._lib.launches - l0
    t_res = ctx.max_over_ranks(e0.elapsed_time(e1) / 1e3)
    ctx.barrier()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(a.steps):
        model.set_input(batch, phase="test")
        model.test()
        out_host.copy_(model.fake_N, non_blocking=True)
        torch.cuda.synchronize()
    t_e2e = ctx.max_over_ranks(time.perf_counter() - t0)
    n = ctx.world_size
    line = {"metric": "skitG generator forward images/sec", "value": n * B * a.steps / t_res, "unit": UNIT, "n_gpus": n, "steps": a.steps,
            "warmup": max(a.warmup, 3), "ms_per_step": 1e3 * t_res / a.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "bf16x3 split operands, fp32 accumulate; fp32 elsewhere", "data": "synthetic",
            "config": {"workload": "generator forward %dx%d, batch %d per rank, resnet_9blocks ngf64 (BASELINE.json configs[4])" % (size, size, B),
                       "size": size, "batch_per_rank": B, "parallelism": "replicas x%d (no collective)" % n,
                       "l2": "activations per forward exceed the 126 MB L2"},
            "e2e": {"value": n * B * a.steps / t_e2e, "unit": UNIT, "h2d_bytes_per_step": int(model.h2d_bytes), "d2h_bytes_per_step": int(out_host.numel() * 4)},
            "gpu_launches": int(launches), "clocks": clk.summary()}
    if ctx.rank == 0:
        print(json.dumps(line), flush=True)
    if n > 1:
        torch.cuda.synchronize()
        ctx.barrier()
        sys.stdout.flush()
        os._exit(0)


def main():
    a = parse()
    if a.impl == "reference":
        run_reference(a)
    elif a.impl == "eager":
        run_eager(a)
    elif a.mode == "infer":
        run_infer(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
