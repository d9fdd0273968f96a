"""One weight-gradient shape on the tcgen05 kernel, a few launches (for `ncu -k regex:k_wgrad_tc`).
Usage: python tools/wgrad_one.py N Cin Cout D H W KD KH KW"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as entry  # noqa: E402

entry.build()
from megaportrait_hack_b200 import ops  # noqa: E402

N, Ci, Co, D, H, W, KD, KH, KW = (int(a) for a in sys.argv[1:10])
x = ops._to_cl_act(torch.randn(N, Ci, D, H, W, device="cuda"))
g = ops._to_cl_act(torch.randn(N, Co, D, H, W, device="cuda"))
ops.ensure_split(x), ops.ensure_split(g)
for _ in range(4):
    ops.conv_weight_grad(x, g, (KD, KH, KW))
torch.cuda.synchronize()
