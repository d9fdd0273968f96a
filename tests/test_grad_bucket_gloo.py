"""Host logic of the data-parallel training step (engine.GradBucket / DataParallelTrainer wiring): world_size 2 over gloo on CPU
with a stand-in module (the collective and the bucket views are device-agnostic; the CUDA path is tests/test_gpu_train.py)."""
import os
import socket
import sys

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, out_dir, n_buckets):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from megaportrait_hack_b200 import engine
    torch.manual_seed(0)                                   # same initial weights on every rank
    net = torch.nn.Sequential(torch.nn.Linear(6, 5), torch.nn.Tanh(), torch.nn.Linear(5, 3))
    bucket = engine.GradBucket(net, n_buckets=n_buckets)
    assert len(bucket.ranges) == n_buckets and bucket.ranges[0][0] == 0 and bucket.ranges[-1][1] == bucket.flat.numel()
    assert sum(c for _, _, c in bucket.ranges) == len(bucket.params)
    opt = torch.optim.SGD(bucket.params, lr=0.1)
    g = torch.Generator().manual_seed(100 + rank)          # different data per rank
    x = torch.randn(4, 6, generator=g)
    for _ in range(2):
        bucket.zero()
        # this rank's own gradient, computed without touching .grad (with n_buckets > 1 the hooks reduce during backward)
        own = torch.autograd.grad(net(x).pow(2).mean(), bucket.params)
        local = torch.zeros_like(bucket.flat)
        for gp, p, off in zip(own, bucket.params, bucket.offsets):
            local[off:off + p.numel()] = gp.reshape(-1)
        net(x).pow(2).mean().backward()
        if n_buckets > 1:                                  # every range was sent by its last gradient's hook
            assert all(bucket._launched)
        bucket.all_reduce_mean()
        opt.step()
    for p, off in zip(bucket.params, bucket.offsets):      # grads are still views of the flat buffer after the steps
        assert p.grad.data_ptr() == bucket.flat[off:].data_ptr()
    torch.save((rank, local, bucket.flat.clone(), [p.detach().clone() for p in net.parameters()]),
               os.path.join(out_dir, f"r{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


import pytest


@pytest.mark.parametrize("n_buckets", [1, 3])
def test_grad_bucket_all_reduce_world2(tmp_path, n_buckets):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_worker, args=(2, port, str(tmp_path), n_buckets), nprocs=2, join=True)
    res = [torch.load(os.path.join(tmp_path, f"r{r}.pt")) for r in range(2)]
    (_, l0, f0, w0), (_, l1, f1, w1) = res
    assert not torch.equal(l0, l1)                                   # ranks saw different data
    assert torch.allclose(f0, (l0 + l1) / 2, atol=1e-7) and torch.equal(f0, f1)      # one averaged gradient everywhere
    for a, b in zip(w0, w1):
        assert torch.equal(a, b)                                     # weights stay identical across ranks


class _ToyG(torch.nn.Module):
    """Stand-in with Gbase's calling convention: forward(xs, xd) -> (image, pyramids)."""

    def __init__(self):
        super().__init__()
        self.a = torch.nn.Conv2d(3, 4, 3, padding=1)
        self.b = torch.nn.Conv2d(4, 3, 3, padding=1)

    def forward(self, xs, xd):
        y = torch.sigmoid(self.b(torch.relu(self.a(xs + xd))))
        return y, {"prediction_0.5": torch.nn.functional.avg_pool2d(y, 2)}


def _trainer_worker(rank, world, port, out_dir):
    sys.path.insert(0, ROOT)
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from megaportrait_hack_b200 import engine
    torch.manual_seed(0)
    net = _ToyG()
    tr = engine.DataParallelTrainer(net, lambda ps: torch.optim.AdamW(ps, lr=1e-2), n_buckets=2,
                                    loss_fn=lambda pred, pyr, xs, xd: (pred - xd).abs().mean() + pyr["prediction_0.5"].mean())
    g = torch.Generator().manual_seed(10 + rank)
    xs, xd = torch.rand(2, 3, 8, 8, generator=g), torch.rand(2, 3, 8, 8, generator=g)
    losses = [float(tr.step(xs, xd)) for _ in range(4)]
    torch.save((losses, [p.detach().clone() for p in net.parameters()]), os.path.join(out_dir, f"t{rank}.pt"))
    dist.barrier()
    dist.destroy_process_group()


def test_data_parallel_trainer_world2(tmp_path):
    """engine.DataParallelTrainer (eager): two ranks with different data keep bit-identical weights and their losses fall."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    mp.spawn(_trainer_worker, args=(2, port, str(tmp_path)), nprocs=2, join=True)
    (l0, w0), (l1, w1) = [torch.load(os.path.join(tmp_path, f"t{r}.pt")) for r in range(2)]
    assert l0 != l1 and l0[-1] < l0[0] and l1[-1] < l1[0]
    for a, b in zip(w0, w1):
        assert torch.equal(a, b)
