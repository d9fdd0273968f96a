#!/usr/bin/env python
"""The fused RGB stem (conv3x3 + BN + ReLU + max-pool) of the motion-encoder trunks at batch 32: ms per launch."""
import sys, math, torch
import os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from megaportrait_hack_b200 import lib, ops
lib.build()
x = torch.rand(32, 3, 512, 512, device="cuda")
w = torch.randn(128, 3, 3, 3) / math.sqrt(27)
pw = ops.pack_stem3x3_f16(w, torch.randn(128) * 0.1, "cuda")
for _ in range(3):
    out = ops.stem3x3_relu_maxpool_f16(x, pw)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(20):
    out = ops.stem3x3_relu_maxpool_f16(x, pw)
e1.record(); torch.cuda.synchronize()
print("fused stem ms", e0.elapsed_time(e1) / 20)
