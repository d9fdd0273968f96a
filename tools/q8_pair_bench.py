#!/usr/bin/env python
"""A/B of the fp16 + FP8 512 -> 512 3x3 convolution @64x64 x 32 frames (G2d res-blocks): CTA-pair kernel (cta_group::2) vs
the single-CTA kernel (MPB200_TC_PAIR=0).  CUDA events, L2 flushed between launches, best / median of 10."""
import math
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from megaportrait_hack_b200 import lib, ops  # noqa: E402


def main():
    lib.build()
    dev = "cuda"
    N, C, H, W = 32, 512, 64, 64
    g = torch.Generator().manual_seed(0)
    x = torch.randn(N, 1, H, W, C, generator=g)
    a = ops.Act(tuple(x.shape), h16=x.half().to(dev), q8=ops.q8_planes(x, 16.0).to(dev), q8_scale=16.0)
    w = torch.randn(C, C, 3, 3, generator=g) / math.sqrt(C * 9)
    pw = ops.pack_conv(w, torch.zeros(C), dev, prec=ops.PREC_F16_Q8)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    flops = 2 * N * H * W * C * C * 9
    for mode in ("1", "0", "1", "0"):
        os.environ["MPB200_TC_PAIR"] = mode
        for _ in range(3):
            ops.conv(a, pw, res=a, act=ops.ACT_RELU, f32=False, hq=True, out_q8_scale=16.0)
        ts = []
        for _ in range(10):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            ops.conv(a, pw, res=a, act=ops.ACT_RELU, f32=False, hq=True, out_q8_scale=16.0)
            e1.record()
            torch.cuda.synchronize()
            ts.append(e0.elapsed_time(e1))
        print(f"MPB200_TC_PAIR={mode}: best {min(ts):.4f} ms ({flops / min(ts) / 1e9:.1f} useful TFLOP/s), median "
              f"{statistics.median(ts):.4f} ms ({flops / statistics.median(ts) / 1e9:.1f})", flush=True)


if __name__ == "__main__":
    main()
