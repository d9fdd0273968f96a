#!/usr/bin/env python
"""In-graph time of every stage of the source half (one CUDA graph per stage, replayed): developer tool, run on a B200 via gpurun."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__)))); sys.path.insert(0, os.path.join(sys.path[0], "tests"))
import __graft_entry__ as entry
entry.build()
from conftest import synthetic_pair
from megaportrait_hack_b200 import ops
G, sd = entry.load_seeded_gbase("cuda")
xs, xd = synthetic_pair(1)
xs = xs.cuda()
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
def graphed(fn):
    fn(); fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    l0 = ops.LAUNCHES
    with torch.cuda.graph(g):
        out = fn()
    return g, ops.LAUNCHES - l0, out
with torch.no_grad():
    src = G.encode_source(xs, keep_stages=True)
    parts = {
        "Eapp.volume": lambda: G.appearanceEncoder._volume_cl(xs),
        "Eapp.descriptor": lambda: G.appearanceEncoder._descriptor(xs),
        "Emtn(source)": lambda: G._emtn(xs),
        "S2C": lambda: G.warp_generator_s2c._em_theta(src["Rs"], src["ts"], src["zs"], src["es"]),
        "warp": lambda: ops.warp_fused(src["vs"], src["em_s2c"], src["theta_s2c"], sum_d=False, f32=True, split=True),
        "G3d": lambda: G.G3d._forward_cl(src["vc"]),
    }
    for name, fn in parts.items():
        g, n, _ = graphed(fn)
        print("%-16s in-graph %.3f ms  launches %d" % (name, timed(g.replay), n), flush=True)
