#!/usr/bin/env python
"""grid_sample op benchmark of SURVEY.md 8d: v ~ N(0,1) seed 2, (B, 96, 16, 64, 64), B in {1, 32}; grids
reference-faithful / spread / adversarial; implementations brick / brick without bank buckets / workspace / direct.
CUDA events on the launching stream, L2 flushed (a 256 MB write) before every timed launch, best and median of `--reps`.

    python tools/gs_bench.py [--reps 10] [--batches 1,32] [--tune tz,ty,tx,BD,BH,BW,threads,groups] [--json out.json]
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

ALG_BYTES = 51118080


def make_grid(kind: str, B: int, dev):
    import torch
    D, H, W = 16, 64, 64
    zz, yy, xx = torch.meshgrid(torch.linspace(-1, 1, D), torch.linspace(-1, 1, H), torch.linspace(-1, 1, W), indexing="ij")
    base = torch.stack((xx, yy, zz), -1)[None]
    if kind == "spread":
        g = torch.Generator().manual_seed(3)
        return (base + (torch.rand(B, D, H, W, 3, generator=g) - 0.5) * 0.2).to(dev)
    if kind == "adversarial":
        g = torch.Generator().manual_seed(4)
        return ((torch.rand(B, D, H, W, 3, generator=g) - 0.5) * 3.0).to(dev)
    if kind == "reference":
        # what a10 -> a11 hands to F.grid_sample: g' = 2 (id + flow) / (size - 1) - 1 with
        # id = linspace(-1, 1) and flow in [0, 1) + a rigid part: everything lands in the first cells of the volume
        g = torch.Generator().manual_seed(5)
        flow = torch.rand(B, D, H, W, 3, generator=g)
        size = torch.tensor([W - 1.0, H - 1.0, D - 1.0])
        return (2.0 * (base + flow) / size - 1.0).to(dev)
    raise ValueError(kind)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=10)
    ap.add_argument("--batches", default="1,32")
    ap.add_argument("--impls", default="brick,brick_nobucket,ws,direct")
    ap.add_argument("--grids", default="reference,spread,adversarial")
    ap.add_argument("--tune", default="")
    ap.add_argument("--json", default="")
    ap.add_argument("--check", action="store_true", help="compare every implementation with the direct kernel")
    ap.add_argument("--no-flush", action="store_true")
    args = ap.parse_args()
    import torch
    from megaportrait_hack_b200 import lib, ops
    lib.build()
    dev = torch.device("cuda")
    peak = 6460.9
    try:
        peak = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
    except Exception:
        pass
    if args.tune:
        ops.gs_brick_tune([int(x) for x in args.tune.split(",")])
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    rows = []
    for B in [int(b) for b in args.batches.split(",")]:
        v = torch.randn(B, 96, 16, 64, 64, generator=torch.Generator().manual_seed(2)).to(dev)
        for kind in args.grids.split(","):
            grid = make_grid(kind, B, dev)
            want = ops.grid_sample3d(v, grid, impl="direct") if args.check else None
            for impl in args.impls.split(","):
                kw = {"brick_nobucket": dict(impl="brick", bucket=False),
                      "brick_inline": dict(impl="brick", second_pass=False)}.get(impl, dict(impl=impl))
                if impl == "copy":       # same bytes through torch's copy kernel + a grid-sized read: what this size can reach
                    out = torch.empty_like(v)
                    run = lambda: (out.copy_(v), grid.sum())
                else:
                    run = lambda: ops.grid_sample3d(v, grid, **kw)
                if impl == "copy":
                    run()
                else:
                    out = run()
                if want is not None and impl != "copy":
                    assert torch.equal(out, want), (B, kind, impl)
                for _ in range(2):
                    run()
                ts = []
                for _ in range(args.reps):
                    if not args.no_flush:
                        flush.fill_(1)
                    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                    e0.record()
                    run()
                    e1.record()
                    torch.cuda.synchronize()
                    ts.append(e0.elapsed_time(e1))
                best, med = min(ts), statistics.median(ts)
                nbytes = B * ALG_BYTES
                row = {"batch": B, "grid": kind, "impl": impl, "best_ms": best, "median_ms": med,
                       "gbs_best": nbytes / best / 1e6, "gbs_median": nbytes / med / 1e6,
                       "frac_best": nbytes / best / 1e6 / peak, "frac_median": nbytes / med / 1e6 / peak}
                rows.append(row)
                print(f"B={B:3d} {kind:11s} {impl:15s} best {best:8.4f} ms  median {med:8.4f} ms  "
                      f"{row['gbs_median']:8.1f} GB/s (median)  frac {row['frac_median']:.3f}  best frac {row['frac_best']:.3f}",
                      flush=True)
            del grid
        del v
    if args.json:
        json.dump({"peak_hbm_gbs": peak, "l2_flush": not args.no_flush, "tune": args.tune, "rows": rows},
                  open(args.json, "w"), indent=1)


if __name__ == "__main__":
    main()
