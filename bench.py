#!/usr/bin/env python
"""bench.py -- Gbase 512x512 driver frames/sec (BASELINE.json metric) on N B200s of one node.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--drivers-per-gpu 32]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

One "step" = one source frame encoded (Eapp, Emtn, S2C warp, G3d) on rank 0, ONE NCCL broadcast of the encoded
source volume + descriptor (25.2 MB) when N > 1, then `drivers-per-gpu` driver frames per rank through
Emtn, C2D warp generator, fused warp + depth sum, G2d and the image pyramid (BASELINE config 2 at N = 1, config 3's
per-GPU share at N > 1: weak scaling, 32 driver frames per GPU).  Nothing is cached across steps.

Beside the headline the JSON line carries: `e2e` (pinned host buffers in and out every step), `roofline` + `kernels`
(CUDA events around every launch of one eager step), `grid_sample` (the op benchmark of SURVEY.md 8d: B in {1, 32} x
three grids), `parity_vs_n1` (per-frame checksums of every rank against one GPU), `strong_scaling` (BASELINE config 3:
256 frames over the N GPUs), `config2A` (reference semantics `Gbase.forward(xs.expand(32), xd)`), `config1` (B = 1
latency) and `cpu_baseline`.

`--impl reference` times the reference algorithm on the host CPU (the oracle port of model.py; the Python reference
itself does not travel to the GPU box) on a bounded sample of the same workload and prints the same `config`.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

METRIC = "gbase_512x512_driver_frames_per_sec"
UNIT = "frames/s"
FLOPS_PER_DRIVER = 505.4e9     # per driver frame, source cached (hot-path rows, excl. Emtn)
FLOPS_SOURCE = 1065.0e9        # Eapp + S2C + G3d per source


def peaks():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(path):
        d = json.load(open(path))
        return {"hbm_gbs": d["hbm_gbs"], "tf_burst": d["bf16_tflops"], "tf_sustained": d["bf16_tflops_sustained"],
                "src": "measured"}
    return {"hbm_gbs": 6650.0, "tf_burst": 1590.0, "tf_sustained": 1400.0, "src": "fallback"}


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.idx = gpu_index
        self.f = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.p = None

    def start(self):
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "200", "-i", str(self.idx)], stdout=self.f, stderr=subprocess.DEVNULL)
        except Exception:
            self.p = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.p is None:
            return out
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.f.read().splitlines():
            c = [x.strip() for x in line.split(",")]
            if len(c) < 9:
                continue
            try:
                sm.append(float(c[1])); mx.append(float(c[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), c[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(mx), reasons=sorted(reasons), samples=len(sm))
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return out


def synthetic_inputs(n_drivers_total: int):
    import torch
    g = torch.Generator().manual_seed(1)
    xs = torch.rand(1, 3, 512, 512, generator=g)
    xd = torch.rand(n_drivers_total, 3, 512, 512, generator=g)
    return xs, xd


GS_ALG_BYTES = 51118080       # (2 * 96 * 16 * 64 * 64 + 3 * 16 * 64 * 64) * 4 per sample


def gs_grid(kind: str, B: int, dev):
    """The three grids of SURVEY.md 8d for a (16, 64, 64) volume."""
    import torch
    D, H, W = 16, 64, 64
    zz, yy, xx = torch.meshgrid(torch.linspace(-1, 1, D), torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    base = torch.stack((xx, yy, zz), -1)[None]
    if kind == "spread":            # identity + U(-0.1, 0.1), seed 3
        return (base + (torch.rand(B, D, H, W, 3, generator=torch.Generator().manual_seed(3)) - 0.5) * 0.2).to(dev)
    if kind == "adversarial":       # U(-1.5, 1.5), seed 4
        return ((torch.rand(B, D, H, W, 3, generator=torch.Generator().manual_seed(4)) - 0.5) * 3.0).to(dev)
    # reference-faithful: what a10 -> a11 hands to F.grid_sample: 2 (id + flow) / (size - 1) - 1 with
    # id = linspace(-1, 1) and flow in [0, 1): every voxel samples the first cells of the volume
    flow = torch.rand(B, D, H, W, 3, generator=torch.Generator().manual_seed(5))
    return (2.0 * (base + flow) / torch.tensor([W - 1.0, H - 1.0, D - 1.0]) - 1.0).to(dev)


def grid_sample_leg(dev, pk, batches=(1, 32), reps=7):
    import torch
    from megaportrait_hack_b200 import ops
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    def timed(fn):
        for _ in range(2):
            fn()
        ts = []
        for _ in range(reps):
            flush.fill_(1)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a0.record()
            fn()
            a1.record()
            torch.cuda.synchronize()
            ts.append(a0.elapsed_time(a1))
        return statistics.median(ts)

    cells = []
    with torch.no_grad():
        for B in batches:
            v = torch.randn(B, 96, 16, 64, 64, generator=torch.Generator().manual_seed(2)).to(dev)
            out = torch.empty_like(v)
            nbytes = B * GS_ALG_BYTES
            for kind in ("reference", "spread", "adversarial"):
                grid = gs_grid(kind, B, dev)
                # random-permutation grids: nothing fits a brick, the channels-last workspace gather is the right tool
                impl = "ws" if kind == "adversarial" else "brick"
                ms = timed(lambda: ops.grid_sample3d(v, grid, impl=impl))
                ms_copy = timed(lambda: (out.copy_(v), grid.sum()))
                gbs = nbytes / ms / 1e6
                cells.append({"batch": B, "grid": kind, "impl": impl, "ms": ms, "achieved": gbs, "unit": "GB/s",
                              "frac": gbs / pk["hbm_gbs"], "copy_same_bytes_ms": ms_copy,
                              "frac_of_copy": ms_copy / ms})
                del grid
            del v, out
    head = next(c for c in cells if c["batch"] == 32 and c["grid"] == "spread")
    traffic = None
    try:
        t = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))["k_grid_sample_brick|32x96x16x64x64 spread"]
        traffic = {"dram_bytes_per_launch": t["dram_bytes"], "algorithmic_bytes_per_launch": t["algorithmic_bytes"],
                   "ratio": t["dram_bytes"] / t["algorithmic_bytes"], "source": t["source"]}
    except Exception:
        pass
    return {"op": "mp_grid_sample3d_brick (TMA-staged bricks; NCDHW in/out, no workspace copy) / mp_grid_sample3d_ws for "
                  "the adversarial grid", "bound": "hbm", "peak": pk["hbm_gbs"], "unit": "GB/s",
            "algorithmic_bytes_per_sample": GS_ALG_BYTES, "l2": "flushed before every timed launch", "timing": f"median of {reps}",
            "batch": head["batch"], "grid": head["grid"], "ms": head["ms"], "achieved": head["achieved"], "frac": head["frac"],
            "traffic": traffic, "cells": cells}


# ------------------------------------------------------------------------------------------------- reference arm
def workload_config(drivers_per_gpu: int, world: int) -> dict:
    """The `config` object of BOTH arms (the driver compares them): the workload and nothing that depends on the
    implementation.  One step = one source frame + `drivers_per_gpu` driver frames per GPU; the source-only half (Eapp,
    Emtn(source), S2C warp, G3d) is evaluated once per step and shared by the step's driver frames -- SURVEY.md 8d
    config 2 variant (B); variant (A), everything recomputed per pair, is reported beside it under `config2A` /
    `per_pair`."""
    return {"workload": f"Gbase inference, 1 src x {drivers_per_gpu} drv per GPU, 512x512 (BASELINE config "
                        f"{'2' if world == 1 else '3, weak-scaling share'}); source half evaluated once per step, nothing "
                        "cached across steps",
            "drivers_per_gpu": drivers_per_gpu, "global_drivers": drivers_per_gpu * world,
            "parallelism": f"driver-shard x{world}",
            "collective": "none" if world == 1 else "1 broadcast of vc2d+es (25.2 MB) per step",
            "weights": "seeded synthetic (megaportrait_hack_b200/seeded.py, seed 0)",
            "inputs": "torch.rand, generator seed 1 (SURVEY.md 8d)"}


CPU_SAMPLE_DRIVERS = 2


def cpu_step_sample(sd, xs, xd):
    """Bounded sample of one step on the host CPU (oracle = the reference's algorithm): the source half ONCE and the
    per-driver half for CPU_SAMPLE_DRIVERS frames.  Returns (seconds for the source half, seconds per driver frame)."""
    import torch
    import gbase_oracle as O
    k = CPU_SAMPLE_DRIVERS
    with torch.no_grad():
        t0 = time.perf_counter()
        src = O.encode_source(xs, sd)
        t1 = time.perf_counter()
        srcN = {kk: (v.expand(k, *v.shape[1:]) if torch.is_tensor(v) else v) for kk, v in src.items()}
        O.drive(srcN, xd[:k].contiguous(), sd)
        t2 = time.perf_counter()
    return t1 - t0, (t2 - t1) / k


def cpu_measure(sd, warmup: int, steps: int, drivers_per_step: int):
    """-> dict: frames/s of the `drivers_per_step`-driver step extrapolated from the bounded samples (the per-driver half is
    linear in the number of frames: eval mode has no cross-sample coupling), and the reference-semantics per-pair rate."""
    xs, xd = synthetic_inputs(CPU_SAMPLE_DRIVERS)
    for _ in range(warmup):
        cpu_step_sample(sd, xs, xd)
    ts, td = [], []
    for _ in range(steps):
        a, b = cpu_step_sample(sd, xs, xd)
        ts.append(a); td.append(b)
    t_src, t_drv = sum(ts) / len(ts), sum(td) / len(td)
    step_s = t_src + drivers_per_step * t_drv
    return {"value": drivers_per_step / step_s, "step_s_extrapolated": step_s, "source_half_s": t_src,
            "per_driver_frame_s": t_drv, "per_pair_frames_per_s": 1.0 / (t_src + t_drv),
            "sample_s_total": sum(ts) + CPU_SAMPLE_DRIVERS * sum(td)}


def cpu_sample_note(drivers_per_step: int, warmup: int, steps: int) -> str:
    return (f"oracle/gbase_oracle.py (CPU restatement of reference model.py, ATen fp32, all host threads): {warmup} warm-up + "
            f"{steps} timed samples, each = the source half once + the per-driver half for {CPU_SAMPLE_DRIVERS} frames; "
            f"frames/s of the {drivers_per_step}-driver step = {drivers_per_step} / (t_source + {drivers_per_step} * "
            f"t_per_driver_frame)")


def run_reference(args):
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from megaportrait_hack_b200 import seeded
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = seeded.seeded_state_dict(seed=0)
    B = args.drivers_per_gpu
    m = cpu_measure(sd, args.warmup, args.steps, B)
    line = {
        "impl": "reference", "metric": METRIC, "value": m["value"], "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * m["step_s_extrapolated"], "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": workload_config(B, args.gpus),
        "cpu_baseline": {"value": m["value"], "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": cpu_sample_note(B, args.warmup, args.steps)},
        "e2e": {"value": m["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
        "per_pair": {"value": m["per_pair_frames_per_s"], "unit": UNIT,
                     "note": "reference semantics Gbase(xs, xd) with Bs == Bd: source half recomputed for every pair "
                             "(SURVEY.md 8d config 2 variant A)"},
        "source_half_s": m["source_half_s"], "per_driver_frame_s": m["per_driver_frame_s"],
        "cpu_sample_s_per_step": m["sample_s_total"] / max(args.steps, 1),
    }
    print(json.dumps(line), flush=True)


# ------------------------------------------------------------------------------------------------- B200 arm
def run_b200(args):
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    from megaportrait_hack_b200 import lib, ops

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    if rank == 0:
        entry.build()
    if world > 1:
        dist.barrier()
    lib.load()
    G, sd = entry.load_seeded_gbase(dev)
    B = args.drivers_per_gpu
    xs_h, xd_all = synthetic_inputs(B * world)
    xs_h = xs_h.pin_memory()
    xd_h = xd_all[rank * B:(rank + 1) * B].contiguous().pin_memory()
    if rank != 0 or world == 1:
        xd_all = None          # rank 0 keeps every rank's frames for the parity-vs-one-GPU leg
    xs_d, xd_d = xs_h.to(dev), xd_h.to(dev)
    rgb_h = torch.empty((B, 3, 512, 512), dtype=torch.float32).pin_memory()
    from megaportrait_hack_b200.engine import GraphedGbase, ShardedGbase
    # rank 0 encodes; the path's only collective is one 25.2 MB broadcast of vc2d + es
    graphed, graph_note = None, "off (--no-graphs)"
    if not args.no_graphs:
        try:
            graphed = GraphedGbase(G, B, dev)
            graph_note = "on: {encode_source || motion encoder} and render replayed from two CUDA graphs"
        except Exception as e:   # capture is an optimisation; the eager path is the same kernels
            graphed, graph_note = None, f"off (capture failed: {type(e).__name__}: {str(e)[:120]})"
            torch.cuda.synchronize()
    if world > 1:
        flag = torch.tensor([1.0 if graphed is not None else 0.0], device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN)
        if flag.item() < 0.5:
            graphed = None
    sharded = ShardedGbase(G)

    def step(xs, xd):
        if graphed is not None:
            return graphed.step(xs, xd)
        return sharded.step(xs, xd)

    def sync_all():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
            torch.cuda.synchronize()

    if args.profile_step:
        # `ncu --profile-from-start off ...`: warm up, then expose exactly ONE eager step to the profiler and exit
        with torch.no_grad():
            for _ in range(max(args.warmup, 1)):
                sharded.step(xs_d, xd_d)
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStart()
            sharded.step(xs_d, xd_d)
            torch.cuda.synchronize()
            torch.cuda.cudart().cudaProfilerStop()
        if world > 1:
            dist.destroy_process_group()
        return
    with torch.no_grad():
        for _ in range(args.warmup):
            step(xs_d, xd_d)
        # ---- device-resident timing
        sampler = ClockSampler(local)
        sync_all()
        if rank == 0:
            sampler.start()
        l0 = ops.LAUNCHES
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            step(xs_d, xd_d)
        e1.record()
        sync_all()
        launches = ops.LAUNCHES - l0
        ms = e0.elapsed_time(e1)
        # ---- end-to-end: pinned host inputs -> device -> result back on the host, EVERY step, all inside the timed
        # region.  Copies run on their own streams and are double-buffered, so the H2D of step i+1 and the D2H of
        # step i-1 overlap the kernels of step i (what a serving loop does); nothing is skipped or cached.
        sync_all()
        cur = torch.cuda.current_stream()
        s_in, s_out = torch.cuda.Stream(), torch.cuda.Stream()
        xs_buf = [torch.empty_like(xs_d) for _ in range(2)]
        xd_buf = [torch.empty_like(xd_d) for _ in range(2)]
        in_ready = [torch.cuda.Event() for _ in range(2)]
        step_done = [torch.cuda.Event() for _ in range(2)]
        f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        t_host0 = time.perf_counter()
        f0.record()
        s_in.wait_event(f0)

        def upload(i):
            with torch.cuda.stream(s_in):
                if i >= 2:
                    s_in.wait_event(step_done[i % 2])       # buffer i%2 was read by step i-2
                xs_buf[i % 2].copy_(xs_h, non_blocking=True)
                xd_buf[i % 2].copy_(xd_h, non_blocking=True)
                in_ready[i % 2].record(s_in)

        upload(0)
        for i in range(args.steps):
            if i + 1 < args.steps:
                upload(i + 1)
            cur.wait_event(in_ready[i % 2])
            rgb, _ = step(xs_buf[i % 2], xd_buf[i % 2])
            if graphed is not None:
                rgb = rgb.clone()        # graph outputs are static; free them for the next replay before the D2H runs
            step_done[i % 2].record(cur)
            rgb.record_stream(s_out)
            with torch.cuda.stream(s_out):
                s_out.wait_event(step_done[i % 2])
                rgb_h.copy_(rgb, non_blocking=True)
        host_enqueue_ms = (time.perf_counter() - t_host0) * 1e3 / args.steps
        cur.wait_stream(s_out)
        f1.record()
        sync_all()
        clocks = sampler.stop() if rank == 0 else None
        ms_e2e = f0.elapsed_time(f1)
        # ---- roofline leg: one extra instrumented step, CUDA events around every hot launch on its own stream
        sharded.step(xs_d, xd_d)          # untimed eager step: fills the (non-graph) allocator pool
        torch.cuda.synchronize()
        ops.PROFILE = []
        sharded.step(xs_d, xd_d)          # eager, so that every launch can be bracketed by events
        torch.cuda.synchronize()
        prof, ops.PROFILE = ops.PROFILE, None

    # ---- parity at the measured configuration: per-frame checksums (float64 sum of the RGB frame) of one more step on
    # every rank, against rank 0 driving the same frames alone (same kernels => the sums agree to the last bit)
    parity = None
    with torch.no_grad():
        rgb, _ = step(xs_d, xd_d)
        fp = rgb.double().sum(dim=(1, 2, 3))
        if world > 1:
            allfp = [torch.empty_like(fp) for _ in range(world)]
            dist.all_gather(allfp, fp)
            if rank == 0:
                worst = 0.0
                for r in range(world):
                    xr = xd_all[r * B:(r + 1) * B].to(dev)
                    ref, _ = G.drive(G.encode_source(xs_d), xr)
                    worst = max(worst, ((ref.double().sum(dim=(1, 2, 3)) - allfp[r]).abs() / allfp[r].abs()).max().item())
                    del xr, ref
                parity = {"checked_frames": B * world, "max_rel_checksum_diff": worst, "ok": bool(worst <= 1e-9),
                          "what": "per-frame float64 sum of RGB on every rank vs the same frame driven by rank 0 alone"}
            dist.barrier()
        else:
            ref, _ = G.drive(G.encode_source(xs_d), xd_d)
            worst = ((ref.double().sum(dim=(1, 2, 3)) - fp).abs() / fp.abs()).max().item()
            parity = {"checked_frames": B, "max_rel_checksum_diff": worst, "ok": bool(worst <= 1e-9),
                      "what": "per-frame float64 sum of RGB, graph replay vs eager Gbase.drive(encode_source(xs), xd)"}
            del ref
        torch.cuda.synchronize()

    # ---- strong scaling (BASELINE config 3): 256 driver frames of ONE source over the N GPUs, 256 / N per GPU in chunks of
    # `drivers_per_gpu`; the source is encoded and broadcast once per step
    strong = None
    if graphed is not None and args.strong_drivers > 0 and (args.strong_drivers % (B * world) == 0):
        per = args.strong_drivers // world
        with torch.no_grad():
            xbig = xd_d.repeat(per // B, 1, 1, 1)
            for _ in range(2):
                graphed.step_chunks(xs_d, xbig)
            sync_all()
            s0, s1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s0.record()
            for _ in range(args.strong_steps):
                graphed.step_chunks(xs_d, xbig)
            s1.record()
            sync_all()
            strong_ms = s0.elapsed_time(s1) / args.strong_steps
            del xbig
        strong = strong_ms
    # ---- BASELINE config 2 variant (A) and config 1 through the drop-in `Gbase.forward` (eager launches, rank 0, N = 1)
    cfgA = cfg1 = cfg4 = cfg5 = None
    if world == 1 and not args.no_extra_configs:
        with torch.no_grad():
            def timed_fwd(xs_in, xd_in, reps):
                for _ in range(3):                 # first call eager, second captures the shape's graph, third replays
                    G(xs_in, xd_in)
                torch.cuda.synchronize()
                f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                f0.record()
                for _ in range(reps):
                    G(xs_in, xd_in)
                f1.record()
                torch.cuda.synchronize()
                return f0.elapsed_time(f1) / reps
            msA = timed_fwd(xs_d.expand(B, -1, -1, -1).contiguous(), xd_d, 3)
            cfgA = {"value": B / (msA * 1e-3), "unit": UNIT, "ms_per_step": msA,
                    "what": f"Gbase.forward(xs.expand({B}), xd): reference semantics, source half recomputed for every pair "
                            "(SURVEY.md 8d config 2 variant A) through the drop-in forward() (which replays a shape-keyed CUDA graph "
                            "from the second call on)"}
            G.release_forward_graphs()             # the batch-32 graph owns ~10 GB of activations: drop it before the next legs
            torch.cuda.empty_cache()
            ms1 = timed_fwd(xs_d, xd_d[:1].contiguous(), 5)
            G.forward_graphs = False
            ms1_eager = timed_fwd(xs_d, xd_d[:1].contiguous(), 5)
            G.forward_graphs = True
            cfg1 = {"latency_ms_forward": ms1, "latency_ms_forward_eager_launches": ms1_eager,
                    "what": "BASELINE config 1 on B200: the drop-in Gbase.forward(xs, xd) with 1 source + 1 driver frame, device-resident "
                            "inputs, outputs cloned out of the graph's static buffers; `eager_launches` = forward_graphs off"}
            # BASELINE config 4 (row f-3): the high-resolution stage.  Genh is fully convolutional (model.py:1349-1391):
            # timed on a batch of 8 frames at 1024 x 1024, and GHR = Genh(Gbase(xs, xd)) on 8 (src, drv) pairs at 512 x 512
            from megaportrait_hack_b200 import model as M, seeded as S
            genh = M.Genh().eval()
            genh.load_state_dict(S.genh_state_dict(0))
            genh = genh.to(dev)
            xg = torch.rand(8, 3, 1024, 1024, generator=torch.Generator().manual_seed(11)).to(dev) * 2 - 1

            def timed(fn, reps):
                for _ in range(3):                 # (GHR goes through Gbase.forward: eager, capture, replay)
                    fn()
                torch.cuda.synchronize()
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                for _ in range(reps):
                    fn()
                t1.record()
                torch.cuda.synchronize()
                return t0.elapsed_time(t1) / reps
            ms_g = timed(lambda: genh(xg), 3)
            hw = 1024 * 1024
            blk = lambda px: 2 * 2 * px * 64 * 64 * 9            # one ResBlock2D(64, 64): two 3x3 convs
            fl = 2 * hw * 64 * 3 * 49 + blk(hw) + blk(hw // 4) + blk(hw // 16) + 9 * blk(hw // 64) + blk(hw // 16) + \
                blk(hw // 4) + blk(hw) + 2 * hw * 3 * 64 * 49
            ghr = M.GHR().eval()
            ghr.Gbase, ghr.Genh = G, genh
            ms_h = timed(lambda: ghr(xs_d.expand(8, -1, -1, -1).contiguous(), xd_d[:8].contiguous()), 3)
            cfg4 = {"genh_1024x1024_batch8": {"value": 8 / (ms_g * 1e-3), "unit": UNIT, "ms": ms_g,
                                              "useful_tflops": 8 * fl / (ms_g * 1e-3) / 1e12},
                    "ghr_512x512_8_pairs": {"value": 8 / (ms_h * 1e-3), "unit": UNIT, "ms": ms_h},
                    "what": "BASELINE config 4: Genh (model.py:1349-1391, ResBlock2D(64) read as ResBlock2D(64, 64)) on 8 frames "
                            "of 1024 x 1024, and GHR.forward = Genh(Gbase(xs, xd)[0]) on 8 pairs; eager launches, seeded weights"}
            del genh, ghr, xg
            # BASELINE config 5, generator half (row f-2): one `train_base` iteration of Gbase alone -- train-mode forward
            # (batch-statistics BatchNorm), an L1 loss against the driver frame in place of the out-of-scope perceptual /
            # adversarial losses, backward through every operator of the path on libmpb200, AdamW step (train.py:135, 194, 318).
            # Runs on a second model instance so the headline model keeps its eval-mode weights.
            cfg5 = None
            if not args.no_train_leg:
                import __graft_entry__ as entry
                Gt = entry.load_seeded_gbase(dev)[0].train()       # a second instance: the headline model keeps its eval-mode state
                opt = torch.optim.AdamW(Gt.parameters(), lr=1e-5, betas=(0.5, 0.999), weight_decay=1e-2)
                xs1, xd1 = xs_d[:1].contiguous(), xd_d[:1].contiguous()

                def train_iter():
                    with torch.enable_grad():            # (this block of legs runs under torch.no_grad())
                        opt.zero_grad(set_to_none=True)
                        pred, _ = Gt(xs1, xd1)
                        loss = (pred - xd1).abs().mean()
                        loss.backward()
                    opt.step()
                    return loss.detach()
                n0 = ops.LAUNCHES
                train_iter()
                launches_train = ops.LAUNCHES - n0
                torch.cuda.synchronize()
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                for _ in range(2):
                    loss = train_iter()
                t1.record()
                torch.cuda.synchronize()
                ms_t = t0.elapsed_time(t1) / 2
                # the same iteration replayed as ONE CUDA graph (whole-network capture: forward, backward and a capturable AdamW):
                # at batch 1 the eager iteration is bound by the host enqueueing ~7 000 launches
                ms_graph, graph_note_t = None, None
                try:
                    opt = torch.optim.AdamW(Gt.parameters(), lr=1e-5, betas=(0.5, 0.999), weight_decay=1e-2, capturable=True, fused=True)
                    side = torch.cuda.Stream()
                    side.wait_stream(torch.cuda.current_stream())
                    with torch.cuda.stream(side):
                        for _ in range(2):
                            train_iter()
                    torch.cuda.current_stream().wait_stream(side)
                    torch.cuda.synchronize()
                    gt = torch.cuda.CUDAGraph()
                    opt.zero_grad(set_to_none=True)
                    with torch.cuda.graph(gt):
                        with torch.enable_grad():
                            pred, _ = Gt(xs1, xd1)
                            loss_g = (pred - xd1).abs().mean()
                            loss_g.backward()
                        opt.step()
                    gt.replay()
                    torch.cuda.synchronize()
                    t0.record()
                    for _ in range(3):
                        gt.replay()
                    t1.record()
                    torch.cuda.synchronize()
                    ms_graph = t0.elapsed_time(t1) / 3
                    graph_note_t = f"loss after replay {float(loss_g):.6f}"
                    del gt
                except Exception as e:  # noqa: BLE001  (the eager number above stands on its own)
                    graph_note_t = f"capture failed: {type(e).__name__}: {str(e)[:200]}"
                    torch.cuda.synchronize()
                # the same iteration at batch 4 (4 different (source, driver) pairs): the launches carry 4x the work, so the
                # iteration is no longer bound by launch latency / the host
                ms_b4 = None
                try:
                    nb = 4
                    xs4 = torch.cat([xs_d[:1], xd_d[4:7]], 0).contiguous()       # four distinct source frames
                    xd4 = xd_d[:nb].contiguous()
                    # through engine.DataParallelTrainer (one process: no collective), the iteration replayed as one CUDA graph
                    # so that the number does not depend on the box's host speed
                    from megaportrait_hack_b200 import engine as E
                    tr4 = E.DataParallelTrainer(
                        Gt, lambda ps: torch.optim.AdamW(ps, lr=1e-5, betas=(0.5, 0.999), weight_decay=1e-2, capturable=True, fused=True),
                        graph=not args.no_graphs, warmup=1)
                    for _ in range(3):                      # eager, capture, first replay
                        tr4.step(xs4, xd4)
                    torch.cuda.synchronize()
                    t0.record()
                    for _ in range(2):
                        tr4.step(xs4, xd4)
                    t1.record()
                    torch.cuda.synchronize()
                    del tr4
                    ms_b4 = t0.elapsed_time(t1) / 2
                except Exception as e:  # noqa: BLE001
                    graph_note_t = (graph_note_t or "") + f"; batch-4 leg failed: {type(e).__name__}: {str(e)[:160]}"
                    torch.cuda.synchronize()
                # 2041 GF forward per pair; backward = data + weight gradient of every convolution ~ 2x the forward
                cfg5 = {"generator_fwd_bwd_adamw_batch1": {"ms": ms_t, "pairs_per_s": 1e3 / ms_t, "loss": float(loss),
                                                           "ms_cuda_graph": ms_graph, "cuda_graph": graph_note_t,
                                                           "libmpb200_launches": launches_train,
                                                           "useful_tflops": 3 * 2041e9 / ((ms_graph or ms_t) * 1e-3) / 1e12,
                                                           "batch4_ms": ms_b4,
                                                           "batch4_pairs_per_s": None if not ms_b4 else 4e3 / ms_b4,
                                                           "batch4_useful_tflops": None if not ms_b4 else
                                                           4 * 3 * 2041e9 / (ms_b4 * 1e-3) / 1e12},
                        "what": "BASELINE config 5, generator half at batch 1 on one GPU: Gbase.train() forward + backward + AdamW "
                                "with an L1 loss (losses / discriminator are out of scope); fp32-grade three-pass convolutions in "
                                "all three directions (weight gradient on tcgen05), operators exchange channels-last views; `ms` = eager "
                                "launches with the reference's default AdamW, `ms_cuda_graph` / `batch4_ms` = the whole iteration "
                                "replayed as one CUDA graph with AdamW(capturable, fused)"}
                del Gt, opt

    # ---- BASELINE config 5 across the GPUs (row f-2): data-parallel training of the generator, one (source, driver) pair per rank,
    # ONE NCCL all-reduce of the flat gradient buffer per iteration (engine.DataParallelTrainer); weak scaling, max over ranks
    cfg5_dp = None
    if world > 1 and not args.no_train_leg and not args.no_extra_configs:
        from megaportrait_hack_b200 import engine as E
        Gt = entry.load_seeded_gbase(dev)[0]
        trainer = E.DataParallelTrainer(
            Gt, lambda ps: torch.optim.AdamW(ps, lr=1e-5, betas=(0.5, 0.999), weight_decay=1e-2, capturable=True, fused=True),
            graph=not args.no_graphs, warmup=2,       # the iteration incl. the NCCL all-reduce is replayed as one CUDA graph
            n_buckets=args.train_buckets)             # > 1: bucket all-reduces overlap the rest of the backward pass
        xs1, xd1 = xs_d[:1].contiguous(), xd_d[:1].contiguous()          # every rank owns different driver frames
        for _ in range(4):                                               # 2 eager + capture + 1 replay
            loss = trainer.step(xs1, xd1)
        sync_all()
        d0, d1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        d0.record()
        for _ in range(3):
            loss = trainer.step(xs1, xd1)
        d1.record()
        sync_all()
        tms = torch.tensor([d0.elapsed_time(d1) / 3], dtype=torch.float64, device=dev)
        dist.all_reduce(tms, op=dist.ReduceOp.MAX)
        # the replicas must still hold identical weights: compare a float64 checksum of all parameters across ranks
        chk = torch.stack([p.detach().double().sum() for p in trainer.bucket.params]).sum().reshape(1)
        lo, hi = chk.clone(), chk.clone()
        dist.all_reduce(lo, op=dist.ReduceOp.MIN)
        dist.all_reduce(hi, op=dist.ReduceOp.MAX)
        cfg5_dp = {"ms_per_iteration": float(tms.item()), "pairs_per_s": world * 1e3 / float(tms.item()), "scaling": "weak",
                   "pairs_per_gpu": 1, "all_reduce_bytes": int(trainer.bucket.flat.numel() * 4), "loss_rank0": float(loss),
                   "all_reduce_buckets": len(trainer.bucket.ranges),
                   "replica_checksum_spread": float((hi - lo).abs().item()),
                   "what": "BASELINE config 5, generator half, data-parallel: per rank Gbase.train() forward + backward on its own "
                           "pair, one NCCL all-reduce of the flat fp32 gradient buffer, AdamW on every rank; the whole iteration incl. the "
                           "collective replayed as one CUDA graph (--no-graphs: eager launches)"}
        del Gt, trainer
    t = torch.tensor([ms, ms_e2e, strong if strong is not None else 0.0], dtype=torch.float64, device=dev)
    cnt = torch.tensor([float(launches)], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
    ms, ms_e2e, strong_ms = t.tolist()
    frames = B * world * args.steps
    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    pk = peaks()
    agg = {}
    if args.dump_launches:
        with open(args.dump_launches, "w") as f:
            for kind, a, b, fl, by in prof:
                if kind.startswith("stage:"):
                    continue
                t_ms = a.elapsed_time(b)
                f.write(f"{t_ms:9.4f} ms  {fl / t_ms / 1e9 if fl else 0:8.1f} TFLOP/s  {by / t_ms / 1e6 if by else 0:8.1f} GB/s  {kind}\n")
    stages = {}
    for kind, a, b, fl, by in prof:
        if kind.startswith("stage:"):
            t_ms = a.elapsed_time(b)
            stages[kind[6:]] = {"ms": round(t_ms, 3),
                                "useful_tflops": round(fl / (t_ms * 1e-3) / 1e12, 1) if fl else None}
    prof = [x for x in prof if not x[0].startswith("stage:")]
    for kind, a, b, fl, by in prof:
        kind = kind.split("|")[0]
        d = agg.setdefault(kind, {"ms": 0.0, "flops": 0, "bytes": 0, "n": 0})
        d["ms"] += a.elapsed_time(b); d["flops"] += fl; d["bytes"] += by; d["n"] += 1
    step_ms = ms / args.steps
    roof = None
    traffic = None
    try:
        tj = json.load(open(os.path.join(ROOT, "profiles", "traffic.json")))
        t = tj.get("k_conv_q8_pair|32x1x64x64 512->512 k133 s1 (MP_PREC_F16_Q8, cta_group::2)") or \
            tj.get("k_conv_tc2|32x1x64x64 512->512 k133 s1 (MP_PREC_F16_Q8)") or tj.get("k_conv_tc2|32x1x64x64 512->512 k133 s1")
        if t:
            traffic = {"dram_bytes_per_launch": t["dram_bytes"], "algorithmic_bytes_per_launch": t["algorithmic_bytes"],
                       "launch": "G2d 512->512 3x3 @64x64 x32 (16 of the step's conv launches)", "source": t["source"]}
    except Exception:
        traffic = None
    fam = [(k, agg[k], passes) for k, passes in (("conv_tc", 3), ("conv_tc_h", 2), ("conv_tc_q8", 2)) if k in agg]
    if fam:
        # the dominant kernel family: k_conv_tc2 / k_conv_tc3 in its three operand modes (pass-units per product: 3 for
        # split-bf16, 2 for fp16x2, 2 for fp16 + FP8 cross terms: one fp16 pass + two FP8 passes at twice the rate)
        ms_all = sum(c["ms"] for _, c, _ in fam)
        fl_all = sum(c["flops"] for _, c, _ in fam)
        raw_all = sum(c["flops"] * ps for _, c, ps in fam)
        n_all = sum(c["n"] for _, c, _ in fam)
        ach = fl_all / (ms_all * 1e-3) / 1e12
        family = {"kernel": "k_conv_tc2 / k_conv_tc3 / k_conv_q8_pair (tcgen05 implicit-GEMM conv; split-bf16 x3, fp16 x2 and "
                            "fp16 + e4m3 cross-term operand modes, the last one on cta_group::2)",
                  "achieved": ach, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                  "mma_passes": raw_all / fl_all, "raw_tensor_frac": raw_all / (ms_all * 1e-3) / 1e12 / pk["tf_sustained"],
                  "launches": n_all, "avg_launch_ms": ms_all / n_all, "share_of_step": ms_all / step_ms,
                  "algorithmic_gflop_per_launch": fl_all / n_all / 1e9,
                  "modes": {k: {"launches": c["n"], "ms": c["ms"], "useful_tflops": c["flops"] / (c["ms"] * 1e-3) / 1e12,
                                "pass_units": ps} for k, c, ps in fam}}
        # `roofline` = the DOMINANT KERNEL of the step: k_conv_q8_pair (the fp16 + FP8 cross-term convolutions of G2d on CTA
        # pairs: the largest single share of the step in the ncu launch list, profiles/round2_*_launches_bench.txt);
        # ALGORITHMIC flops (2 * M * N * K of the convolution, once -- not the 2 pass-units the operand format costs) over the
        # summed launch durations.  `family` = the same figure over every tensor-core convolution launch of the step.
        dom = agg.get("conv_tc_q8")
        if dom and dom["ms"] > 0:
            ach_d = dom["flops"] / (dom["ms"] * 1e-3) / 1e12
            roof = {"bound": "tensor", "kernel": "k_conv_q8_pair (tcgen05 cta_group::2 implicit-GEMM convolution, MP_PREC_F16_Q8: fp16 "
                                                 "main product + e4m3 cross terms; G2d res-blocks and up-blocks 1-2)",
                    "achieved": ach_d, "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach_d / pk["tf_sustained"],
                    "traffic": traffic, "peak_source": pk["src"] + " bf16 sustained (kernel timed inside a long step)",
                    "mma_passes": 2, "raw_tensor_frac": 2 * ach_d / pk["tf_sustained"],
                    "launches": dom["n"], "avg_launch_ms": dom["ms"] / dom["n"], "share_of_step": dom["ms"] / step_ms,
                    "algorithmic_gflop_per_launch": dom["flops"] / dom["n"] / 1e9,
                    "note": "pass-units per product: one fp16 pass + two FP8 passes at twice the rate = 2; ncu: tensor pipe 87.9 % "
                            "active (profiles/round2_q8_pair_full.txt)",
                    "family": family}
        else:                     # (A/B runs with the FP8 format switched off: report the family)
            roof = dict(family, bound="tensor", traffic=traffic,
                        peak_source=pk["src"] + " bf16 sustained (kernel timed inside a long step)")
    extra = {}
    if "conv_tc_h" in agg:     # two-pass fp16 convolutions of the motion-encoder trunks
        c = agg["conv_tc_h"]
        ach = c["flops"] / (c["ms"] * 1e-3) / 1e12
        extra["conv_tc_f16x2"] = {"kernel": "k_conv_tc2/3, MP_PREC_F16X2 (Emtn trunks)", "bound": "tensor", "achieved": ach,
                                  "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                                  "mma_passes": 2, "raw_tensor_frac": 2 * ach / pk["tf_sustained"], "launches": c["n"],
                                  "ms": c["ms"], "share_of_step": c["ms"] / step_ms}
    if "conv_tc_q8" in agg:    # fp16 main product + FP8 cross terms (G2d identity res-blocks)
        c = agg["conv_tc_q8"]
        ach = c["flops"] / (c["ms"] * 1e-3) / 1e12
        extra["conv_tc_f16_q8"] = {"kernel": "k_conv_q8_pair (cta_group::2), MP_PREC_F16_Q8 (G2d res-blocks)", "bound": "tensor", "achieved": ach,
                                   "peak": pk["tf_sustained"], "unit": "TFLOP/s", "frac": ach / pk["tf_sustained"],
                                   "mma_passes": 2, "raw_tensor_frac": 2 * ach / pk["tf_sustained"], "launches": c["n"],
                                   "ms": c["ms"], "share_of_step": c["ms"] / step_ms,
                                   "note": "pass-units: one fp16 pass + two FP8 passes at twice the rate"}
    for kind in ("warp_fused_sum", "warp_fused", "conv_simt"):
        if kind in agg:
            c = agg[kind]
            e = {"launches": c["n"], "ms": c["ms"], "share_of_step": c["ms"] / step_ms}
            if c["bytes"]:
                gbs = c["bytes"] / (c["ms"] * 1e-3) / 1e9
                e.update(bound="hbm", achieved=gbs, peak=pk["hbm_gbs"], unit="GB/s", frac=gbs / pk["hbm_gbs"])
            if c["flops"]:
                e.update(tflops=c["flops"] / (c["ms"] * 1e-3) / 1e12)
            extra[kind] = e

    # ---- grid_sample op leg (the second half of BASELINE.json's metric): F.grid_sample(v, grid, 'bilinear',
    # 'border', align_corners=True), v ~ N(0,1) seed 2 of shape (B, 96, 16, 64, 64), B in {1, 32}, the three grids of 8d,
    # through the reference-layout C-ABI entry points; CUDA events on the launching stream, L2 flushed (256 MB write)
    # before every timed launch, median of 7.  "copy" = the same bytes through torch's copy kernel + a read of the grid:
    # what a launch of this size can reach under the same harness (at B = 1 a 51 MB launch is latency-bound).
    gs = None
    try:
        gs = grid_sample_leg(dev, pk)
    except Exception as e:      # the op leg must never take the headline number down with it
        gs = {"error": f"{type(e).__name__}: {str(e)[:160]}"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        torch.set_num_threads(cores)
        m = cpu_measure(sd, 1, 3, B)        # 1 warm-up + 3 timed
        cpu = {"value": m["value"], "unit": UNIT, "cores": cores, "kind": "port", "sample": cpu_sample_note(B, 1, 3),
               "per_pair_frames_per_s": m["per_pair_frames_per_s"], "source_half_s": m["source_half_s"],
               "per_driver_frame_s": m["per_driver_frame_s"]}

    line = {
        "metric": METRIC, "value": frames / (ms * 1e-3), "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "bf16x3 split (fp32-grade products: hi*hi+hi*lo+lo*hi, fp32 accumulate in TMEM); G2d res-blocks: fp16 "
                                    "main product + e4m3 cross terms; motion-encoder trunks: fp16x2 (fp16 activations x fp16 "
                                    "hi+lo weights, fp32 accumulate)",
        "data": "synthetic",
        "config": workload_config(B, world),
        "engine": {"cuda_graphs": graph_note,
                   "l2": "per-step working set (>10 GB of activations) far exceeds the 126 MB L2; no flush needed"},
        "e2e": {"value": frames / (ms_e2e * 1e-3), "unit": UNIT,
                "h2d_bytes_per_step": int((xs_h.numel() + xd_h.numel()) * 4 * world),
                "d2h_bytes_per_step": int(rgb_h.numel() * 4 * world)},
        "gpu_launches": int(cnt.item()),
        "host_enqueue_ms_per_step": host_enqueue_ms,
        "clocks": clocks,
        "roofline": roof,
        "kernels": extra,
        "grid_sample": gs,
        "parity_vs_n1": parity,
        "strong_scaling": None if not strong_ms else {
            "value": args.strong_drivers / (strong_ms * 1e-3), "unit": UNIT, "ms_per_step": strong_ms, "scaling": "strong",
            "global_drivers": args.strong_drivers, "drivers_per_gpu": args.strong_drivers // world,
            "what": "BASELINE config 3: 1 source x 256 driver frames over the N GPUs (256 / N per GPU in chunks of "
                    f"{B}), source encoded + broadcast once per step; device-resident inputs, max over ranks"},
        "config2A": cfgA,
        "config1": cfg1,
        "config4": cfg4,
        "config5": cfg5,
        "config5_data_parallel": cfg5_dp,
        "stages_eager_step": stages,
        "cpu_baseline": cpu,
        "useful_tflops_whole_step": (B * FLOPS_PER_DRIVER + FLOPS_SOURCE) * world / (step_ms * 1e-3) / 1e12,
    }
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--drivers-per-gpu", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-extra-configs", action="store_true", help="skip the config 2(A) / config 1 legs")
    ap.add_argument("--train-buckets", type=int, default=4,
                    help="gradient all-reduce buckets of the data-parallel training leg (1 = one collective after the backward pass)")
    ap.add_argument("--no-train-leg", action="store_true", help="skip the config 5 leg (Gbase forward + backward + AdamW)")
    ap.add_argument("--strong-drivers", type=int, default=256, help="global driver frames of the strong-scaling leg (0 = off)")
    ap.add_argument("--strong-steps", type=int, default=3)
    ap.add_argument("--no-graphs", action="store_true", help="launch every kernel eagerly instead of replaying CUDA graphs")
    ap.add_argument("--dump-launches", default="", help="write the per-launch CUDA-event profile of one step here")
    ap.add_argument("--profile-step", action="store_true",
                    help="warm up, bracket ONE eager step with cudaProfilerStart/Stop and exit (ncu --profile-from-start off)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
