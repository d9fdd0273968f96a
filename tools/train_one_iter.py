"""Expose exactly ONE training iteration (Gbase.train() forward + backward + AdamW, batch 1) to a profiler:
    ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file out.csv python tools/train_one_iter.py
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as entry  # noqa: E402

entry.build()
G = entry.load_seeded_gbase("cuda")[0].train()
opt = torch.optim.AdamW(G.parameters(), lr=1e-5, betas=(0.5, 0.999), weight_decay=1e-2)
g = torch.Generator().manual_seed(1)
xs = torch.rand(1, 3, 512, 512, generator=g).cuda()
xd = torch.rand(1, 3, 512, 512, generator=g).cuda()


def it():
    opt.zero_grad(set_to_none=True)
    pred, _ = G(xs, xd)
    (pred - xd).abs().mean().backward()
    opt.step()


for _ in range(2):
    it()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStart()
it()
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
