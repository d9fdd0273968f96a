#!/usr/bin/env python
"""mp_conv_wgrad (row f-2) on the path's convolution shapes: ms and useful TFLOP/s (2 * P * Cout * Cin * taps)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megaportrait_hack_b200 import lib, ops  # noqa: E402

lib.build()
for (N, Ci, Co, sp, k) in ((1, 96, 96, (16, 64, 64), (3, 3, 3)), (8, 96, 96, (16, 64, 64), (3, 3, 3)), (8, 512, 512, (1, 64, 64), (1, 3, 3)),
                           (8, 128, 128, (1, 256, 256), (1, 3, 3))):
    x = ops.from_nchw(torch.randn(N, Ci, *sp, device="cuda"), f32=True, split=False)
    g = ops.from_nchw(torch.randn(N, Co, *sp, device="cuda"), f32=True, split=False)
    for _ in range(2):
        ops.conv_weight_grad(x, g, k)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        ops.conv_weight_grad(x, g, k)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    P = N * sp[0] * sp[1] * sp[2]
    fl = 2.0 * P * Ci * Co * k[0] * k[1] * k[2]
    print(f"wgrad N={N} {Ci}->{Co} {sp} k{k}: {ms:.3f} ms  {fl / ms / 1e9:.1f} useful TFLOP/s")
