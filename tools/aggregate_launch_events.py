#!/usr/bin/env python
"""Aggregate `bench.py --dump-launches FILE` (CUDA events around every launch of one eager step) by launch kind.

    python tools/aggregate_launch_events.py gpurun_out/launch_events.txt profiles/roundN_step_launch_profile_events.txt
"""
import collections
import re
import sys


def main():
    src, dst = sys.argv[1], sys.argv[2]
    agg = collections.OrderedDict()
    total = 0.0
    for line in open(src):
        m = re.match(r"\s*([0-9.]+) ms\s+([0-9.]+) TFLOP/s\s+([0-9.]+) GB/s\s+(.*)$", line.rstrip("\n"))
        if not m:
            continue
        ms, tf, gbs, kind = float(m.group(1)), float(m.group(2)), float(m.group(3)), m.group(4).strip()
        d = agg.setdefault(kind, [0.0, 0, 0.0, 0.0])
        d[0] += ms; d[1] += 1; d[2] = tf; d[3] = gbs
        total += ms
    with open(dst, "w") as f:
        f.write(f"# CUDA events around every launch of one eager step (bench.py --dump-launches), aggregated by launch kind; "
                f"total {total:.2f} ms\n")
        for kind, (ms, n, tf, gbs) in sorted(agg.items(), key=lambda kv: -kv[1][0]):
            f.write(f"{ms:8.3f} ms  n={n:3d}  {tf:7.1f} TF/s(last)  {kind}\n")
    print(f"{len(agg)} kinds, {total:.2f} ms -> {dst}")


if __name__ == "__main__":
    main()
