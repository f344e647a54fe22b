"""bench.py — skitG/sinskitG train-step throughput on B200 (BASELINE.json metric).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--size S] [--impl reference]

Workload (config.workload): configs[1] of BASELINE.json — one single-material train step
(G forward, D1 step, D2 step, G step with GAN + L1 + patch-L1, three Adam updates) at SxS = 512x512,
NT = 64 touch patches + NF = 32 random fake patches, PatchNCE off, LPIPS / vision-aided off, on the
tensor-core architecture (resnet_9blocks ngf 64, multiscale PatchGAN ndf 64), batch 1 per rank
(the reference forces batch_size = 1).  Synthetic seeded inputs, random-init weights.

One JSON line on stdout (rank 0).  `value` = images/s with inputs resident in HBM; `e2e` = the same
step through the public model API with host inputs: set_input (pinned H2D) + optimize_parameters +
get_current_losses (D2H) inside the timed region.  N > 1: one process per GPU (torchrun), one sample
per rank per step, flat-bucket gradient all-reduce over NCCL — weak scaling.
`--impl reference` times the CPU oracle (a restatement of the reference's own PyTorch code path,
pinned to the real reference by tests/golden) on the host cores for the same metric.
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


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=40)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--size", type=int, default=512)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--cpu-sample-size", type=int, default=256, help="image side of the bounded CPU-baseline sample")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--mode", default="train", choices=["train", "infer"], help="infer: generator forward only (BASELINE.json configs[4])")
    ap.add_argument("--batch", type=int, default=1, help="images per rank per step (infer mode)")
    ap.add_argument("--nce", action="store_true", help="PatchNCE on (BASELINE.json configs[2] wiring: use with --size 768)")
    ap.add_argument("--lpips", action="store_true", help="LPIPS-VGG16 terms on with the reference's default weights (lambda_G1_lpips 1, "
                                                         "lambda_G2_lpips 10; random VGG weights: no checkpoint offline)")
    return ap.parse_args()


def workload_config(a):
    return {"workload": "skitG/sinskitG train step, single material, %dx%d, resnet_9blocks ngf64 + multiscale PatchGAN ndf64, "
                        "NT=64 NF=32, GAN+L1+patch-L1, PatchNCE %s, LPIPS %s, VAL off (BASELINE.json configs[%d]%s)"
                        % (a.size, a.size, "on (5 layers, 256 patches, T=0.07)" if a.nce else "off",
                           "on (VGG16, full image + touch patches, random weights)" if a.lpips else "off", 2 if a.nce else 1,
                           " + the reference's default LPIPS terms" if a.lpips else ""),
            "size": a.size, "images_per_rank_per_step": 1, "NT": 64, "NF": 32, "netG": "resnet_9blocks", "ngf": 64,
            "netD": "multiscale", "ndf": 64, "parallelism": "dp%d (flat grad bucket all-reduce, NCCL)" % a.gpus,
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


# ----------------------------------------------------------------------------------------- CPU oracle legs
def oracle_step_time(size, steps, warmup, nt, nf, threads, nce=False, lpips=False):
    """Times the CPU oracle's train step (oracle/skit_oracle.py, pinned to the real reference) — the checker
    run as a baseline, never as the product."""
    from oracle import skit_oracle as O
    import vts_b200
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    opt = argparse.Namespace(gan_mode="nonsaturating")
    G = vts_b200.networks.define_G(9, 5, 64, "resnet_9blocks", "instance", False, "xavier", 0.02, False, False, [], opt)
    D = vts_b200.networks.define_D(4, 64, "multiscale", 3, "batch", "xavier", 0.02, False, 3, [], opt)
    D2 = vts_b200.networks.define_D(7, 64, "multiscale", 3, "batch", "xavier", 0.02, False, 3, [], opt)
    sds = [{k: v.detach().clone() for k, v in n.state_dict().items()} for n in (G, D, D2)]
    cfg = O.StepConfig(netG="resnet_9blocks", batch_size_G2=nt, add_fake_T_sample_size=nf, lambda_NCE=1.0 if nce else 0.0,
                       lambda_G1_lpips=1.0 if lpips else 0.0, lambda_G2_lpips=10.0 if lpips else 0.0)
    sdL = O.lpips_random_state(0) if lpips else None
    batch = O.step_inputs_from_batch(O.synthetic_batch(size, NT=nt, seed=0))
    rs = np.random.RandomState(0)
    nce_sizes = [G.feature_hw(l, size, size) for l in cfg.nce_layers] if nce else []
    times = []
    state = {}
    for i in range(warmup + steps):
        rand = dict(real_b=[0.3], real_s=[0.8], fake_b=[0.6], fake_s=[0.2],
                    fake_ox=rs.randint(0, size - 32, nf).astype(np.int32), fake_oy=rs.randint(0, size - 32, nf).astype(np.int32))
        if nce:
            rand["nce_ids"] = [rs.permutation(h * w)[:min(cfg.num_patches, h * w)] for h, w in nce_sizes]
        t0 = time.perf_counter()
        O.train_step(cfg, sds[0], sds[1], sds[2], state, batch, rand, step=i + 1, sdL=sdL)
        if i >= warmup:
            times.append(time.perf_counter() - t0)
    return float(np.mean(times))


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    s = a.cpu_sample_size
    nt, nf = 16, 8
    t = oracle_step_time(s, max(1, min(a.steps, 3)), min(a.warmup, 1), nt, nf, cores, nce=a.nce, lpips=a.lpips)
    # bounded sample: a step at s x s; conv work scales with pixels, so images/s at the full size is scaled by (s/size)^2
    value = (1.0 / t) * (s * s) / float(a.size * a.size)
    sample = "CPU oracle train step at %dx%d (NT=%d NF=%d), %.2f s/step, scaled by (%d/%d)^2 to the %dx%d workload" % (s, s, nt, nf, t, s, a.size, a.size, a.size)
    line = {"impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": 1000.0 / value, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32",
            "data": "synthetic", "config": workload_config(a),
            "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}
    print(json.dumps(line), flush=True)


# ----------------------------------------------------------------------------------------- the B200 arm
def measure_dominant_kernel(size, peaks):
    """The dominant kernel of the step is the tcgen05 ResnetBlock conv (256->256, 3x3, (S/4)^2 pixels; the same kernel
    runs its input gradient).  Timed alone with CUDA events around a replayed CUDA graph of 24 launches (device time, no
    host launch cost) that rotate over 8 operand/output buffer pairs (> 126 MB in total, so no launch finds its input
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
    traffic = None
    try:   # dram bytes per launch of this kernel from the committed ncu --set full capture (profiles/), when present
        prof = json.load(open(os.path.join(ROOT, "profiles", "r01_conv_tc_halo_fwd%d.json" % size)))
        traffic = prof.get("dram_bytes_per_launch")
    except Exception:
        pass
    return {"bound": "tensor", "kernel": "conv_tc_halo_kernel<256> (ResnetBlock conv3x3 256->256 @%dx%d)" % (s, s),
            "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak,
            "peak_source": "measured (MEASURED_PEAKS.json bf16_tflops, burst)" if "bf16_tflops" in peaks else "fallback 1.59 PFLOP/s",
            "traffic": traffic, "ms_per_launch": ms, "timing": "CUDA events around a replayed graph of %d launches over %d buffer pairs (> L2)" % (iters, nbuf),
            "note": "achieved counts ALGORITHMIC flops; the kernel executes 3 bf16 MMAs per product (hi/lo split, the fp32-parity "
                    "requirement: DESIGN.md 4.1), so frac is capped at 1/3; executed tensor-pipe rate %.0f TFLOP/s = %.2f of peak"
                    % (3 * achieved, 3 * achieved / peak)}


def measure_arch_a(a, ctx):
    """The reference's DEFAULT architecture (unet256_custom ngf 10 + multiscale ndf 8: HBM/latency bound)
    through the same step, reported beside the headline tensor-core architecture."""
    import vts_b200
    from oracle import skit_oracle as O
    opt = vts_b200.default_options(netG="unet256_custom", ngf=10, ndf=8, gpu_ids=[ctx.local_rank])
    torch.manual_seed(0)
    m = vts_b200.SinSKITGModel(opt)
    m.set_input(O.synthetic_batch(a.size, NT=64, seed=ctx.rank))
    for _ in range(4):
        m.optimize_parameters(1)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.steps):
        m.optimize_parameters(1)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.steps
    return {"netG": "unet256_custom", "ngf": 10, "ndf": 8, "ms_per_step": ms, "images_per_s_per_gpu": 1e3 / ms}


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
                                   lambda_G1_lpips=1.0 if a.lpips else 0.0, lambda_G2_lpips=10.0 if a.lpips else 0.0)
    torch.manual_seed(0)
    model = vts_b200.SinSKITGModel(opt, dist_ctx=ctx if ctx.world_size > 1 else None)
    ctx.broadcast_params([model.netG, model.netD, model.netD2])
    batch = O.synthetic_batch(a.size, NT=64, seed=ctx.rank)   # a different (material, augmentation) sample per rank
    for k in ("S", "I", "M", "T_images", "I_masks"):
        batch[k] = batch[k].pin_memory()

    def step_resident():
        model.optimize_parameters(1)

    def step_e2e():
        model.set_input(batch)
        model.optimize_parameters(1)
        return model.get_current_losses()

    model.set_input(batch)
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
    d2h = 4 * (8 + 3 * 64 + 32)
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
        line["roofline"] = measure_dominant_kernel(a.size, peaks)
        line["arch_A_default"] = measure_arch_a(a, ctx)
        if n == 1 and not a.no_cpu_baseline:
            cores = os.cpu_count() or 1
            s, nt, nf = a.cpu_sample_size, 16, 8
            t = oracle_step_time(s, 1, 1, nt, nf, cores, nce=a.nce, lpips=a.lpips)
            v = (1.0 / t) * (s * s) / float(a.size * a.size)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": "CPU oracle train step at %dx%d (NT=%d NF=%d), %.2f s/step, scaled by (%d/%d)^2" % (s, s, nt, nf, t, s, a.size)}
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
    elif a.mode == "infer":
        run_infer(a)
    else:
        run_b200(a)


if __name__ == "__main__":
    main()
