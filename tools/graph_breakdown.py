"""Where the graph-replayed step goes: {source encode || driver motion encoder} graph vs render graph, and the source
encode / motion encoder alone (developer tool; run on a B200 via gpurun)."""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import __graft_entry__ as entry  # noqa: E402
from megaportrait_hack_b200.engine import GraphedGbase, pack_source  # noqa: E402


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return min(ts), sum(ts) / len(ts)


def main():
    dev = torch.device("cuda", 0)
    G, _ = entry.load_seeded_gbase(dev)
    B = 32
    g = torch.Generator().manual_seed(1)
    xs = torch.rand(1, 3, 512, 512, generator=g).to(dev)
    xd = torch.rand(B, 3, 512, 512, generator=g).to(dev)
    with torch.no_grad():
        eng = GraphedGbase(G, B, dev)
        eng.step(xs, xd)
        print("pre graph  {source || motion}: best %.2f ms avg %.2f" % timed(eng.g_pre.replay))
        print("render graph                 : best %.2f ms avg %.2f" % timed(eng.g_render.replay))
        # the two halves of the pre graph on their own
        flat = torch.zeros_like(eng.flat)
        g_src = torch.cuda.CUDAGraph()
        s = torch.cuda.Stream()
        s.wait_stream(torch.cuda.current_stream())
        with torch.cuda.stream(s):
            pack_source(G.encode_source(xs), flat)
        torch.cuda.current_stream().wait_stream(s)
        torch.cuda.synchronize()
        with torch.cuda.graph(g_src):
            pack_source(G.encode_source(eng.xs), flat)
        g_mot = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g_mot):
            G.drive_motion(eng.xd)
        print("source encode graph alone    : best %.2f ms avg %.2f" % timed(g_src.replay))
        print("driver motion graph alone    : best %.2f ms avg %.2f" % timed(g_mot.replay))


if __name__ == "__main__":
    main()
