"""Where one Gbase training iteration (config 5, generator half) spends its GPU time: torch.profiler kernel table of a
train-mode forward + backward + AdamW step at batch 1.  Usage (GPU box): python tools/train_profile.py [out.txt]"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch  # noqa: E402
import __graft_entry__ as entry  # noqa: E402


def main():
    entry.build()
    G, _ = entry.load_seeded_gbase("cuda")
    G.train()
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
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0.record()
    it()
    t1.record()
    torch.cuda.synchronize()
    ms = t0.elapsed_time(t1)
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA, ProfilerActivity.CPU]) as prof:
        it()
        torch.cuda.synchronize()
    rows = [e for e in prof.key_averages() if e.device_time_total > 0 and e.device_type.name == "CUDA"]
    rows.sort(key=lambda e: -e.device_time_total)
    total = sum(e.device_time_total for e in rows)
    lines = [f"one training iteration (forward + backward + AdamW, batch 1): {ms:.1f} ms wall on the stream; "
             f"sum of kernel times {total / 1e3:.1f} ms"]
    for e in rows[:40]:
        lines.append(f"{e.device_time_total / 1e3:9.3f} ms {100 * e.device_time_total / total:5.1f} %  x{e.count:<5d} {e.key[:110]}")
    txt = "\n".join(lines)
    print(txt)
    if len(sys.argv) > 1:
        os.makedirs(os.path.dirname(os.path.abspath(sys.argv[1])), exist_ok=True)
        open(sys.argv[1], "w").write(txt + "\n")


if __name__ == "__main__":
    main()
