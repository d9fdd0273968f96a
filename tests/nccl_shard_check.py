"""Worker of tests/test_gpu_multi.py (launched by torch.distributed.run, one rank per GPU): drives `--drivers` frames of one
source through engine.GraphedGbase / ShardedGbase on real NCCL and has rank 0 compare the gathered shard outputs with the
same frames driven by rank 0 alone.  Prints one JSON line."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--drivers", type=int, default=6)
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import __graft_entry__ as entry
    from megaportrait_hack_b200 import engine
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    G, _sd = entry.load_seeded_gbase(dev)
    g = torch.Generator().manual_seed(1)
    xs = torch.rand(1, 3, 512, 512, generator=g).to(dev)
    xd = torch.rand(args.drivers, 3, 512, 512, generator=g).to(dev)
    lo, hi = engine.shard_range(args.drivers, rank, world)
    res = {}
    with torch.no_grad():
        for name in ("sharded", "graphed"):
            if name == "sharded":
                rgb, pyr = engine.ShardedGbase(G).step(xs, xd[lo:hi])
            else:
                eng = engine.GraphedGbase(G, hi - lo, dev)
                eng.step(xs, xd[lo:hi])
                rgb, pyr = eng.step(xs, xd[lo:hi])          # second replay: static buffers reused
            torch.cuda.synchronize()
            parts = [torch.empty((engine.shard_range(args.drivers, r, world)[1] - engine.shard_range(args.drivers, r, world)[0],
                                  3, 512, 512), device=dev) for r in range(world)]
            dist.all_gather(parts, rgb.contiguous())
            if rank == 0:
                src = G.encode_source(xs)
                full, _ = G.drive(src, xd)
                got = torch.cat(parts, 0)
                same_batch = 0.0         # every shard against rank 0 driving exactly that shard (same batch => same bits)
                for r in range(world):
                    a, b = engine.shard_range(args.drivers, r, world)
                    ref_r, _ = G.drive(src, xd[a:b])
                    same_batch = max(same_batch, (parts[r] - ref_r).abs().max().item())
                res[name] = {"max_abs_vs_same_shard_on_one_gpu": same_batch,
                             "max_abs_vs_full_batch_on_one_gpu": (got - full).abs().max().item(),
                             "swap_visible": (got[0] - full[-1]).abs().max().item()}
    if rank == 0:
        print("NCCL_SHARD_CHECK " + json.dumps({"world": world, "drivers": args.drivers, **res}), flush=True)
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
