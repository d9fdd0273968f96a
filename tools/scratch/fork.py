import os, sys, torch
sys.path.insert(0, "/root/repo"); sys.path.insert(0, "/root/repo/tests")
import __graft_entry__ as entry
entry.build()
from conftest import synthetic_pair
from megaportrait_hack_b200.engine import GraphedGbase
G, sd = entry.load_seeded_gbase("cuda")
xs, xd = synthetic_pair(32)
xs, xd = xs.cuda(), xd.cuda()
def timed(fn, reps=10):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps
with torch.no_grad():
    for overlap in (True, False):
        eng = GraphedGbase(G, 32, "cuda", overlap=overlap)
        print("fork", os.environ.get("MPB200_SOURCE_FORK", "1"), "overlap", overlap, "step ms %.2f" % timed(lambda: eng.step(xs, xd)),
              "pre %.2f" % timed(eng.g_pre.replay), "render %.2f" % timed(eng.g_render.replay), flush=True)
        del eng
    g = torch.cuda.CUDAGraph()
    src = G.encode_source(xs)
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        src = G.encode_source(xs)
    print("encode_source graph alone ms %.2f" % timed(g.replay))
    G.forward_graphs = True
    print("forward B=1 ms %.2f" % timed(lambda: G(xs, xd[:1])))
