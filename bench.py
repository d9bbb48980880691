#!/usr/bin/env python
"""bench.py -- fwd+bwd warp+loss throughput (Gpixels/s) of the xpt-mde hot path on B200.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg3|cfg5]

A "step" is one pass of the hot path over one batch of synthetic snippets: pyramids, inverse
warp of 4 sources at 4 scales, L1 + SSIM + edge-aware smoothness, and the backward to
dL/ddepth_ms, dL/ddisp_ms, dL/dpose (reference model/loss_and_metric/losses.py:26-55 under
train_val.py:78-92).  A pixel is one full-resolution target pixel of one snippet.
Prints ONE JSON line on rank 0.  See DESIGN.md "Measurement" for every field.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "xpt-mde-2021_b200")
for _p in (ROOT, PKG):
    if _p not in sys.path:
        sys.path.insert(0, _p)

import torch  # noqa: E402

METRIC = "fwd+bwd warp+loss Gpixels/s"
UNIT = "Gpixel/s"
WORKLOADS = {
    # name: (B, H, W, scaling) -- BASELINE.json configs[1], [2], [3], [4]; "weak": B snippets per rank,
    # "strong": B is the GLOBAL batch, sharded B / world per rank (configs[3]: 64 over 2/4/8 GPUs, configs[4]: 128)
    "cfg2": (8, 128, 384, "weak"),
    "cfg3": (16, 256, 832, "weak"),
    "cfg4": (64, 128, 384, "strong"),
    "cfg5": (128, 384, 1280, "strong"),
}


def workload_string(name, source_grad=False, adversarial=False):
    """the same string in both arms (ours / reference), so the driver sees one config"""
    B, H, W, scaling = WORKLOADS[name]
    return (f"{name}: B={B}{'/gpu' if scaling == 'weak' else ' global'} {H}x{W} snippet=5 (4 sources) scales=4 LOSS_RIGID_T1 fwd+bwd"
            + (" +dL/dsource" if source_grad else "") + (" ADVERSARIAL inputs (iid depth, large poses)" if adversarial else ""))
N_SRC, N_SCALES = 4, 4
GAMMA = sum(1.0 / (4 ** s) for s in range(N_SCALES))          # 85/64
# algorithmic bytes per full-res target pixel (DESIGN.md "Bytes model")
BYTES_PREP = 48 + 48 * (GAMMA - 1) + 12 + 12 * (GAMMA - 1)     # read frames, write levels s>1: 79.7
BYTES_FUSED = N_SRC * GAMMA * 12 + GAMMA * 12 + 4 * GAMMA * 4   # gather src_ms, tgt_ms, depth, disp, d_depth, d_disp: 100.9
BYTES_SURVEY_STEP = 355.9                                      # unfused fwd + bwd + prep model of SURVEY 8d


def load_peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
    return 6650.0, "fallback (B200_PROFILING.md)"


def bind_to_gpu_numa_node(index):
    """Pin this rank to the CPUs next to its GPU (sysfs local_cpulist of the GPU's PCI function) BEFORE any pinned
    host buffer is allocated, so that the buffers of the end-to-end leg are first-touched on the GPU's NUMA node.
    Returns a short description for the JSON line; silently a no-op when the topology cannot be read."""
    try:
        import pynvml
        pynvml.nvmlInit()
        bus = pynvml.nvmlDeviceGetPciInfo(pynvml.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{dom[-4:].lower()}:{rest.lower()}"
        with open(path + "/local_cpulist") as f:
            spec = f.read().strip()
        cpus = set()
        for part in spec.split(","):
            if "-" in part:
                a, b = part.split("-")
                cpus.update(range(int(a), int(b) + 1))
            elif part:
                cpus.add(int(part))
        cpus &= os.sched_getaffinity(0)
        if not cpus:
            return None
        os.sched_setaffinity(0, cpus)
        node = None
        try:
            with open(path + "/numa_node") as f:
                node = int(f.read().strip())
        except OSError:
            pass
        return f"numa node {node}, {len(cpus)} cpus"
    except Exception:
        return None


class ClockSampler:
    """nvidia-smi clocks / throttle reasons DURING the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index, self.rows, self.proc = index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except OSError:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                mx.append(float(r[1]))
            except (ValueError, IndexError):
                continue
            for n, v in zip(names, r[2:6]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def make_input_sets(orc, B, H, W, n_sets, seed, device, adversarial=False):
    """n_sets distinct snippet batches resident in HBM (rotated so that no step finds its inputs in L2).
    Two are generated from scratch; the rest are cheap perturbations of those (rolled images, scaled
    depth, jittered pose) -- generating 20+ full batches on the CPU would dominate start-up."""
    base = [orc.make_inputs(B, H, W, N=N_SRC, n_scales=N_SCALES, seed=seed + i, adversarial=adversarial) for i in range(2)]
    sets = []
    for i in range(n_sets):
        bf, bp = base[i % 2]
        if i < 2:
            feats, preds = bf, bp
        else:
            g = torch.Generator().manual_seed(seed + i)
            feats = {"image5d": torch.roll(bf["image5d"], shifts=(i, 3 * i), dims=(2, 3)), "intrinsic": bf["intrinsic"]}
            preds = {"depth_ms": [d * (1 + 0.01 * i) for d in bp["depth_ms"]],
                     "disp_ms": [d / (1 + 0.01 * i) for d in bp["disp_ms"]],
                     "pose": bp["pose"] * (1 + 0.02 * torch.rand(bp["pose"].shape, generator=g))}
        sets.append(({k: v.to(device) for k, v in feats.items()},
                     {"depth_ms": [d.to(device) for d in preds["depth_ms"]],
                      "disp_ms": [d.to(device) for d in preds["disp_ms"]], "pose": preds["pose"].to(device)}))
    return sets, base


def set_bytes(feats, preds):
    n = feats["image5d"].numel() + feats["intrinsic"].numel() + preds["pose"].numel()
    n += sum(d.numel() for d in preds["depth_ms"]) + sum(d.numel() for d in preds["disp_ms"])
    return 4 * n


def run_reference(args, B, H, W):
    """--impl reference: the reference's CPU implementation of the path, i.e. the oracle restatement of
    its TF op graph on torch-CPU with all host threads (TF 2.4.1 is not installable offline)."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import xpt_oracle as orc
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    feats, preds = orc.make_inputs(B, H, W, N=N_SRC, n_scales=N_SCALES, seed=20211 + 2000)
    lw, sw = orc.LOSS_RIGID_T1, orc.SCALE_WEIGHT_T1

    def step(fb, pb):
        orc.loss_and_grads(fb, pb, lw, sw, global_batch=B)
    # bound the run: size the per-step sample from one probe step on a single snippet
    sub = lambda b: ({k: v[:b] for k, v in feats.items()},
                     {"depth_ms": [d[:b] for d in preds["depth_ms"]], "disp_ms": [d[:b] for d in preds["disp_ms"]],
                      "pose": preds["pose"][:b]})
    f1, p1 = sub(1)
    step(f1, p1)
    t0 = time.perf_counter()
    step(f1, p1)
    t1 = time.perf_counter() - t0
    budget = 150.0
    sb = int(max(1, min(B, budget / max(1e-6, (args.steps + args.warmup) * t1))))
    fb, pb = sub(sb)
    for _ in range(args.warmup):
        step(fb, pb)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step(fb, pb)
    dt = (time.perf_counter() - t0) / args.steps
    val = sb * H * W / dt / 1e9
    sample = f"{sb} of {B} snippets per step ({H}x{W}, 4 sources, 4 scales, fwd+bwd), {args.steps} steps"
    line = {
        "impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt * 1e3 * (B / sb), "higher_is_better": True,
        "scaling": WORKLOADS[args.workload][3],
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": workload_string(args.workload, args.source_grad, args.adversarial),
                   "note": "reference TF op graph restated on torch-CPU (TF 2.4.1 not installable offline)"},
        "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=sorted(WORKLOADS))
    ap.add_argument("--no-graph", action="store_true", help="eager launches instead of the CUDA-graph replay")
    ap.add_argument("--unfused", action="store_true", help="one kernel per reference stage (A/B)")
    ap.add_argument("--source-grad", action="store_true", help="also produce dL/dsource (config 5's full backward)")
    ap.add_argument("--net-grad-mb", type=float, default=0.0,
                    help="N > 1 only: also all-reduce a stand-in fp32 net-gradient bucket of this many MB every step "
                         "(SURVEY 8e: the depth/pose nets are out of scope; ~120 MB is EfficientNet-B5 + PoseNet), "
                         "asynchronously, overlapped with the next step's kernels")
    ap.add_argument("--adversarial", action="store_true",
                    help="SURVEY 8d's adversarial inputs: iid depth U(1,80) per pixel (worst-case gather locality) and "
                         "large poses (about half of the samples leave the image)")
    ap.add_argument("--strip", action="store_true", help="the streaming strip kernel (XPT_FLAG_STRIP) instead of the tile kernel (A/B)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--e2e-depth", type=int, default=2, help="host-buffer steps kept in flight by the e2e leg (1 = synchronous calls)")
    ap.add_argument("--cpu-steps", type=int, default=10)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    B, H, W, scaling = WORKLOADS[args.workload]
    if args.impl == "reference":
        run_reference(args, B, H, W)
        return

    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl ours needs a CUDA device: xptwarp has no CPU fallback")
    torch.cuda.set_device(local_rank)
    # only multi-rank runs bind (several ranks competing for the host side of PCIe); the 1-GPU run also times the
    # CPU baseline, whose worker threads must keep every host core
    numa = bind_to_gpu_numa_node(local_rank) if (world > 1 and os.environ.get("XPT_NO_NUMA_BIND") is None) else None
    device = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=device)
        # torchrun exports OMP_NUM_THREADS=1; the synthetic inputs are generated on the host
        torch.set_num_threads(max(1, (os.cpu_count() or 1) // world))

    import xptwarp
    from xptwarp import _cabi
    from oracle import xpt_oracle as orc     # input generator + cpu_baseline checker only

    lw, sw = orc.LOSS_RIGID_T1, orc.SCALE_WEIGHT_T1
    if scaling == "strong":                 # B is the global batch: contiguous shards of B / world snippets
        if B % world:
            raise SystemExit(f"{args.workload}: global batch {B} does not divide over {world} ranks")
        global_batch, B = B, B // world
    else:
        global_batch = B * world            # weak scaling: B snippets per rank (compute_average_loss divisor)
    # world > 1: the path's only exchange -- the 4 loss scalars (distributer.py:93-110) -- is an ncclAllReduce the
    # library enqueues on the step's own stream behind the epilogue kernel, i.e. one more node of the step's CUDA graph
    flags = ((0 if args.no_graph else _cabi.XPT_FLAG_GRAPH) | (_cabi.XPT_FLAG_UNFUSED if args.unfused else 0)
             | (_cabi.XPT_FLAG_STRIP if args.strip else 0) | (_cabi.XPT_FLAG_ALLREDUCE if world > 1 else 0))
    plan = xptwarp.get_plan(local_rank, B, N_SRC, H, W, [1, 2, 4, 8], sw, lw["L1"], lw["SSIM"], lw["smoothe"],
                            global_batch, flags)
    exchange = "none (1 GPU)"
    if world > 1:
        plan.comm_init(dist)
        exchange = ("the 4 loss scalars summed INSIDE k_epilogue through peer memory (CUDA IPC inboxes over NVLink), part of "
                    "the step's CUDA graph, no collective launch" if plan.comm_status()[0] else
                    "ncclAllReduce of the 4 loss scalars inside the step's CUDA graph (XPT_FLAG_ALLREDUCE)")

    # ---- inputs: enough distinct resident sets that a step never finds its inputs in the 126 MB L2
    probe_f, probe_p = orc.make_inputs(1, H, W, N=N_SRC, n_scales=N_SCALES, seed=1)
    per_set = set_bytes(probe_f, probe_p) * B
    n_sets = int(min(24, max(3, -(-3 * 126e6 // per_set))))
    sets, sets_cpu = make_input_sets(orc, B, H, W, n_sets, 20211 + 2000 + rank, device, args.adversarial)
    calls = []
    for f, p in sets:
        img = f["image5d"]
        calls.append(plan.bind_total_loss(img[:, :-1], img[:, -1], f["intrinsic"], p["depth_ms"], p["disp_ms"],
                                          p["pose"], want_grad=True, want_source_grad=args.source_grad))
    # a real (non-NULL) stream: the legacy default stream cannot be captured into a CUDA graph
    side = torch.cuda.Stream(device)
    side.wait_stream(torch.cuda.current_stream(device))
    torch.cuda.set_stream(side)
    stream = plan.stream()

    net_grad = None
    net_stream = None
    if dist is not None and args.net_grad_mb > 0:
        net_grad = torch.zeros(int(args.net_grad_mb * 1e6 / 4), dtype=torch.float32, device=device)
        net_stream = torch.cuda.Stream(device)

    def step(i):
        c = calls[i % n_sets]
        c.run(stream)                         # world > 1: includes the in-graph all-reduce of the loss vector
        if net_grad is not None:
            # stand-in for the nets' gradient buckets: the library's NCCL group on a second stream,
            # overlapped with the next step's kernels; one bucket in flight
            net_stream.wait_stream(torch.cuda.current_stream(device))
            with torch.cuda.stream(net_stream):
                plan.allreduce([net_grad])

    def drain():
        if net_stream is not None:
            torch.cuda.current_stream(device).wait_stream(net_stream)

    def barrier():
        drain()
        if dist is not None:
            dist.barrier()
        torch.cuda.synchronize()

    n_warm = max(args.warmup, 2 * n_sets)                # warm-up also captures one graph per input set
    for i in range(n_warm):
        step(i)
    barrier()
    # at least ~0.3 s of the same steps before the clock starts: a step is ~0.1 ms at config 2, and a device coming out
    # of idle measured up to 8 % slow over the first few hundred of them.  The count is fixed from one timing on rank 0
    # and broadcast, so every rank issues the same number of collectives.
    t0 = time.perf_counter()
    for i in range(20):
        step(n_warm + i)
    barrier()
    per = max((time.perf_counter() - t0) / 20, 1e-6)
    extra = torch.tensor([min(5000, int(0.3 / per))], dtype=torch.int64, device=device)
    if dist is not None:
        dist.broadcast(extra, src=0)
    for i in range(int(extra.item())):
        step(n_warm + 20 + i)
    barrier()
    n_warm += 20 + int(extra.item())
    launches_per_step = plan.launches()

    # ---- timed region A: K steps, device-timed on the launching stream ------------------------
    sampler = ClockSampler(local_rank)
    sampler.start()
    time.sleep(0.3)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    e0.record()
    for i in range(args.steps):
        step(i)
    drain()
    e1.record()
    barrier()
    ms_total = e0.elapsed_time(e1)
    if ms_total < 400:      # keep the clock sampler meaningful on tiny workloads: same kernels, untimed.
        # The steps carry the in-graph all-reduce when world > 1, so every rank must issue the SAME number of them:
        # the count comes from rank 0's timing (unmatched all-reduces deadlock).
        fill = torch.tensor([min(20000, int(600.0 / max(ms_total / args.steps, 1e-3)))], dtype=torch.int64, device=device)
        if dist is not None:
            dist.broadcast(fill, src=0)
        for i in range(int(fill.item())):
            calls[i % n_sets].run(stream)
        torch.cuda.synchronize()
    clocks = sampler.stop()
    t = torch.tensor([ms_total], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_step = float(t.item()) / args.steps
    pixels_step = B * H * W * world
    value = pixels_step / (ms_step * 1e-3) / 1e9

    # ---- timed region B: the dominant kernel alone (events around each launch, eager launches) -----
    peak, peak_src = load_peaks()
    kern_ms = None
    pyr_ms = None
    if not args.unfused:
        eager = xptwarp.get_plan(local_rank, B, N_SRC, H, W, [1, 2, 4, 8], sw, lw["L1"], lw["SSIM"], lw["smoothe"],
                                 global_batch, flags & ~(_cabi.XPT_FLAG_GRAPH | _cabi.XPT_FLAG_ALLREDUCE))
        ecalls = []
        for (f, p), c in zip(sets, calls):
            img = f["image5d"]
            ecalls.append(eager.bind_total_loss(img[:, :-1], img[:, -1], f["intrinsic"], p["depth_ms"], p["disp_ms"],
                                                p["pose"], want_grad=True, want_source_grad=args.source_grad, out=c.out))
        for i in range(n_sets):
            ecalls[i].run(stream)
        torch.cuda.synchronize()
        nrec = min(args.steps, 200)
        eager.profile_begin(nrec)
        for i in range(nrec):
            ecalls[i % n_sets].run(stream)
        torch.cuda.synchronize()
        durs = eager.profile_end(nrec)
        kern_ms = sum(durs) / len(durs)
        sd = sorted(durs)
        kern_pct = {"p10": sd[len(sd) // 10], "p50": sd[len(sd) // 2], "p90": sd[(9 * len(sd)) // 10], "n": len(sd)}
        # the step's one HBM-bound kernel, timed the same way: pyramids (+ camera geometry)
        eager.profile_begin(nrec, kernel=1)
        for i in range(nrec):
            ecalls[i % n_sets].run(stream)
        torch.cuda.synchronize()
        pdurs = eager.profile_end(nrec)
        pyr_ms = sum(pdurs) / len(pdurs) if pdurs else None
    px_rank = B * H * W
    roofline = None
    if kern_ms:
        fused_bytes = BYTES_FUSED + (N_SRC * GAMMA * 12 if args.source_grad else 0)
        ach = px_rank * fused_bytes / (kern_ms * 1e-3) / 1e9
        traffic = None
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath) and not args.source_grad:
            with open(tpath) as tf_:
                rec = json.load(tf_).get(args.workload)
            if rec:
                traffic = rec["dram_bytes_per_launch"]        # bytes per launch, from the committed ncu --set full capture
        roofline = {"bound": "hbm", "kernel": "k_strip" if args.strip else "k_fused<grad>", "achieved": ach, "peak": peak, "unit": "GB/s",
                    "frac": ach / peak, "traffic": traffic, "algorithmic_bytes_per_launch": px_rank * fused_bytes,
                    "peak_source": peak_src, "kernel_ms": kern_ms, "kernel_ms_percentiles": kern_pct,
                    "bytes_per_pixel": fused_bytes, "kernel_share_of_step": kern_ms / ms_step if world == 1 else None}
    if roofline is not None and pyr_ms:
        # k_pyramid_tiled: read 60 B (5 frames x 12), write the RGBx source levels (4 x gamma x 16) and the target
        # levels s > 1 ((gamma - 1) x 12) per full-resolution target pixel
        pyr_bytes = 60 + N_SRC * GAMMA * 16 + (GAMMA - 1) * 12
        pach = px_rank * pyr_bytes / (pyr_ms * 1e-3) / 1e9
        roofline["secondary"] = {"bound": "hbm", "kernel": "k_pyramid_tiled" if os.environ.get("XPT_PYRAMID") == "tiled" else "k_pyramid_tma", "achieved": pach, "peak": peak, "unit": "GB/s",
                                 "frac": pach / peak, "kernel_ms": pyr_ms, "bytes_per_pixel": pyr_bytes,
                                 "algorithmic_bytes_per_launch": px_rank * pyr_bytes}
    step_model = {"bytes_per_pixel": BYTES_SURVEY_STEP,
                  "achieved_gbs": value / world * BYTES_SURVEY_STEP, "frac_of_peak": value / world * BYTES_SURVEY_STEP / peak,
                  "note": "SURVEY 8d unfused fwd+bwd+prep bytes model at the measured pixel rate, per GPU"}

    # ---- e2e: same step through the host-buffer C-ABI entry point (H2D + D2H inside) ---------------
    # Two steps in flight, as a training loop with a prefetching input pipeline runs it: two contexts, two streams, two
    # sets of pinned buffers; xpt_total_loss_host_begin(step i+1) is issued before xpt_total_loss_host_end(step i), so one
    # step's host->device copies overlap the other's compute and device->host copies.  Every step still copies ITS inputs
    # from pinned host memory and ITS results back inside the timed region.  The one-call-at-a-time number (the
    # synchronous xpt_total_loss_host) is reported next to it as e2e.sync_value.
    import ctypes as C
    from xptwarp.engine import Plan
    depth = max(1, args.e2e_depth)
    if per_set > 1.5e9:                  # config 5: a second set of pinned buffers (4 GB) and a second ctx (10 GB) buy nothing
        depth = 1                        # at a 90 ms transfer per step
    extra_plans = [Plan(local_rank, B, N_SRC, H, W, [1, 2, 4, 8], sw, lw["L1"], lw["SSIM"], lw["smoothe"], global_batch,
                        flags & ~_cabi.XPT_FLAG_ALLREDUCE) for _ in range(depth - 1)]
    extra_streams = [torch.cuda.Stream(device) for _ in range(depth - 1)]

    def host_side(pl, fcpu, pcpu, st):
        himg = fcpu["image5d"].contiguous().pin_memory()
        hK, hpose = fcpu["intrinsic"].contiguous().pin_memory(), pcpu["pose"].contiguous().pin_memory()
        hdepth = [d.contiguous().pin_memory() for d in pcpu["depth_ms"]]
        hdisp = [d.contiguous().pin_memory() for d in pcpu["disp_ms"]]
        hlosses, hdpose = torch.zeros(4).pin_memory(), torch.zeros(B, N_SRC, 6).pin_memory()
        hdd = [torch.zeros_like(d).pin_memory() for d in hdepth]
        hds = [torch.zeros_like(d).pin_memory() for d in hdepth]
        fr = _cabi.XptFrames()
        fr.source, fr.source_batch_stride, fr.source_frame_stride = himg.data_ptr(), himg.stride(0), himg.stride(1)
        fr.target, fr.target_batch_stride = himg.data_ptr() + N_SRC * himg.stride(1) * 4, himg.stride(0)
        fr.intrinsic = hK.data_ptr()
        o = _cabi.XptLossOutputs()
        o.losses, o.d_pose, o.grad_scale = hlosses.data_ptr(), hdpose.data_ptr(), 1.0
        for s_ in range(N_SCALES):
            o.d_depth_ms[s_], o.d_disp_ms[s_] = hdd[s_].data_ptr(), hds[s_].data_ptr()
        dptr, sptr = _cabi.ptr_array([d.data_ptr() for d in hdepth]), _cabi.ptr_array([d.data_ptr() for d in hdisp])
        nin = 4 * (himg.numel() + hK.numel() + hpose.numel() + sum(d.numel() for d in hdepth) + sum(d.numel() for d in hdisp))
        nout = 4 * (4 + hdpose.numel() + sum(d.numel() for d in hdd) + sum(d.numel() for d in hds))
        return {"plan": pl, "fr": fr, "o": o, "dptr": dptr, "sptr": sptr, "hpose": hpose, "hlosses": hlosses, "stream": st,
                "keep": (himg, hK, hdepth, hdisp, hdpose, hdd, hds), "h2d": nin, "d2h": nout}
    sides = [host_side(plan, *sets_cpu[0], stream)] + [host_side(pl, *sets_cpu[(k + 1) % len(sets_cpu)], st.cuda_stream)
                                                       for k, (pl, st) in enumerate(zip(extra_plans, extra_streams))]
    h2d, d2h = sides[0]["h2d"], sides[0]["d2h"]
    dev_losses = torch.zeros(4, device=device)

    def host_begin(k):
        q = sides[k]
        rc = q["plan"]._lib.xpt_total_loss_host_begin(q["plan"].handle, C.byref(q["fr"]), C.byref(q["dptr"]), C.byref(q["sptr"]),
                                                      q["hpose"].data_ptr(), C.byref(q["o"]), q["stream"])
        if rc != 0:
            _cabi.check(rc)

    def host_end(k):
        q = sides[k]
        rc = q["plan"]._lib.xpt_total_loss_host_end(q["plan"].handle)
        if rc != 0:
            _cabi.check(rc)
        if dist is not None:             # the host entry point's losses are rank-local: sum the 4 floats over the ranks
            dev_losses.copy_(q["hlosses"], non_blocking=True)
            plan.allreduce([dev_losses])
    for i in range(20 * depth):          # a call is captured as a graph on its third use; then let the link settle
        host_begin(i % depth); host_end(i % depth)
    barrier()
    n_e2e = max(10, min(args.steps, 100))
    t0 = time.perf_counter()
    for _ in range(n_e2e):               # one call at a time (returns after its D2H copies have landed)
        host_begin(0); host_end(0)
    torch.cuda.synchronize()
    t_sync = (time.perf_counter() - t0) / n_e2e
    barrier()
    t0 = time.perf_counter()
    for i in range(n_e2e + depth - 1):   # `depth` steps in flight
        if i < n_e2e:
            host_begin(i % depth)
        if i >= depth - 1:
            host_end((i - depth + 1) % depth)
    torch.cuda.synchronize()
    t_e2e = torch.tensor([(time.perf_counter() - t0) / n_e2e, t_sync], dtype=torch.float64, device=device)
    if dist is not None:
        dist.all_reduce(t_e2e, op=dist.ReduceOp.MAX)
    e2e_val = pixels_step / float(t_e2e[0].item()) / 1e9
    e2e_sync = pixels_step / float(t_e2e[1].item()) / 1e9

    # ---- CPU baseline beside it (rank 0, N=1 only; bounded sample) ----------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        fb, pb = sets_cpu[0]
        for _ in range(2):
            ref = orc.loss_and_grads(fb, pb, lw, sw, global_batch=global_batch)
        t0 = time.perf_counter()
        for _ in range(args.cpu_steps):
            orc.loss_and_grads(fb, pb, lw, sw, global_batch=global_batch)
        dt = (time.perf_counter() - t0) / args.cpu_steps
        cpu = {"value": B * H * W / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"{args.cpu_steps} steps of the full batch (B={B}, {H}x{W}) on torch-CPU fp32, 2 warm-up",
               "ms_per_step": dt * 1e3}
        # the same run also checks the GPU result of input set 0 against the oracle
        got = calls[0].run(stream)
        torch.cuda.synchronize()
        rel = abs(float(got["losses"][0]) - float(ref["total"])) / abs(float(ref["total"]))
        cpu["gpu_vs_oracle_total_loss_relerr"] = rel

    if rank == 0:
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_step, "higher_is_better": True,
            "scaling": scaling, "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_string(args.workload, args.source_grad, args.adversarial),
                       "global_batch": global_batch, "batch_per_gpu": B, "parallelism": f"dp{world}",
                       "kernel": "k_strip (XPT_FLAG_STRIP)" if args.strip else "k_fused (tiles)",
                       # --warmup steps as asked, then untimed filler steps of the same kind until the device has been
                       # busy for ~0.3 s (a device coming out of idle measured up to 8 % slow): the total before the clock
                       "untimed_steps_before_clock": n_warm,
                       "l2": f"rotating {n_sets} resident input sets ({n_sets * per_set / 1e6:.0f} MB > 126 MB L2)",
                       "launch": "eager" if args.no_graph else "cuda-graph replay", "fused": not args.unfused,
                       "host_affinity": numa or "unbound",
                       "collectives": exchange + (f" + a {args.net_grad_mb:g} MB stand-in net-gradient bucket per step on a second stream"
                                                  if (world > 1 and args.net_grad_mb > 0) else "")},
            "roofline": roofline, "step_model": step_model, "cpu_baseline": cpu, "clocks": clocks,
            "e2e": {"value": e2e_val, "unit": UNIT, "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "in_flight": depth, "sync_value": e2e_sync,
                    "note": f"xpt_total_loss_host_begin/_end, {depth} steps in flight (as many contexts, streams and pinned buffer sets); "
                            "sync_value = one synchronous xpt_total_loss_host call at a time"},
            "gpu_launches": launches_per_step * args.steps,
        }
        print(json.dumps(line))
    if dist is not None:
        if plan.comm_status()[1]:
            raise SystemExit("loss exchange: a peer's record did not arrive (ranks ran different numbers of steps)")
        dist.barrier()
        plan.comm_destroy()              # collective teardown while the process group is alive, not at interpreter exit
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
