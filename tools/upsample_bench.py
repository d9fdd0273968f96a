#!/usr/bin/env python
"""Bilinear x2 upsample kernels (split-bf16 and fp16 + FP8 outputs) at the three G2d shapes, batch 32: ms and TB/s."""
import sys, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megaportrait_hack_b200 import lib, ops
lib.build()
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
for (C, H) in ((512, 64), (256, 128), (128, 256)):
    x = torch.randn(32, C, H, H, device="cuda")
    a = ops.from_nchw(x, f32=False, split=True)
    del x
    if C > 128:
        ms = timed(lambda: ops.upsample2x_bilinear_hq(a, 16.0))
        by = 32 * H * H * C * 4 + 32 * 4 * H * H * C * 4
        print("hq    C=%d H=%d  %.3f ms  %.2f TB/s" % (C, H, ms, by / ms / 1e9))
    ms = timed(lambda: ops.upsample2x_linear(a, 1, f32=False, split=True))
    by = 32 * H * H * C * 4 + 32 * 4 * H * H * C * 4
    print("split C=%d H=%d  %.3f ms  %.2f TB/s" % (C, H, ms, by / ms / 1e9))
    del a
