"""Driver-frame sharding over the GPUs of one node.

The path shards by driver frame: in eval mode nothing couples two samples, so rank r drives frames
[r*N/P, (r+1)*N/P) of a batch that shares one source.  The only data-path collective is ONE broadcast of the encoded
source state -- the canonical volume `vc2d` (96x16x64x64 fp32 = 25 165 824 B) and the descriptor `es` (2 048 B),
packed into a single buffer -- from the encoding rank.  Outputs stay rank-local.
"""
from __future__ import annotations

from typing import Dict, Optional, Tuple

import torch

from . import ops
from .ops import Act

VOL_SHAPE = (1, 16, 64, 64, 96)          # channels-last [N, D, H, W, C]
VOL_NUMEL = 16 * 64 * 64 * 96
ES_NUMEL = 512
STATE_NUMEL = VOL_NUMEL + ES_NUMEL


def shard_range(n_total: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous, balanced split of n_total driver frames: the first (n_total % world) ranks take one extra."""
    base, extra = divmod(n_total, world)
    lo = rank * base + min(rank, extra)
    return lo, lo + base + (1 if rank < extra else 0)


def pack_source(src: Dict[str, object], out: Optional[torch.Tensor] = None) -> torch.Tensor:
    """{'vc2d': Act (1 sample, fp32), 'es': [1,512]} -> one flat fp32 buffer (the broadcast payload)."""
    vol = src["vc2d"].f32
    es = src["es"]
    if vol.numel() != VOL_NUMEL or es.numel() != ES_NUMEL:
        raise RuntimeError(f"pack_source expects one source sample, got {tuple(vol.shape)} / {tuple(es.shape)}")
    if out is None:
        out = torch.empty(STATE_NUMEL, dtype=torch.float32, device=vol.device)
    out[:VOL_NUMEL].copy_(vol.reshape(-1))
    out[VOL_NUMEL:].copy_(es.reshape(-1))
    return out


def unpack_source(flat: torch.Tensor) -> Dict[str, object]:
    """Views into the flat buffer (no copy)."""
    return {"vc2d": Act(VOL_SHAPE, f32=flat[:VOL_NUMEL].view(VOL_SHAPE)), "es": flat[VOL_NUMEL:].view(1, ES_NUMEL)}


class ShardedGbase:
    """One process per GPU.  `step(xs, xd_local)`: rank `src_rank` encodes the source, one broadcast, every rank drives
    its own frames.  With world size 1 no collective is issued."""

    def __init__(self, gbase, group=None, src_rank: int = 0):
        import torch.distributed as dist
        self.G = gbase
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.group = group
        self.world = self.dist.get_world_size(group) if self.dist else 1
        self.rank = self.dist.get_rank(group) if self.dist else 0
        self.src_rank = src_rank
        self._buf: Optional[torch.Tensor] = None

    def broadcast_source(self, src: Optional[Dict[str, object]], device) -> Dict[str, object]:
        if self.world == 1:
            return src
        if self._buf is None or self._buf.device != torch.device(device):
            self._buf = torch.empty(STATE_NUMEL, dtype=torch.float32, device=device)
        if self.rank == self.src_rank:
            pack_source(src, self._buf)
        self.dist.broadcast(self._buf, self.src_rank, group=self.group)
        return unpack_source(self._buf)

    @torch.no_grad()
    def step(self, xs: torch.Tensor, xd_local: torch.Tensor):
        src = self.G.encode_source(xs) if self.rank == self.src_rank else None
        src = self.broadcast_source(src, xd_local.device)
        if xd_local.shape[0] == 0:          # more ranks than driver frames: this rank took part in the broadcast only
            dev = xd_local.device
            return (torch.empty((0, 3, 512, 512), dtype=torch.float32, device=dev),
                    {"prediction_0.5": torch.empty((0, 3, 256, 256), dtype=torch.float32, device=dev),
                     "prediction_0.25": torch.empty((0, 3, 128, 128), dtype=torch.float32, device=dev)})
        return self.G.drive(src, xd_local)


class GraphedGbase:
    """`ShardedGbase.step` replayed from CUDA graphs (fixed shapes: one 512x512 source, `n_drivers` driver frames per
    rank).  A step launches ~470 kernels whose host-side enqueue costs about as much as their GPU time; replaying
    graphs removes that cost.  Graph 1 runs the source encode (source rank only; batch 1, many small kernels) on a
    forked stream CONCURRENTLY with the motion encoder of the driver frames (independent of the source); the broadcast
    of the packed source state stays an eager NCCL call; graph 2 renders.  Inputs are copied into static buffers,
    outputs are static tensors that the next `step` overwrites."""

    def __init__(self, gbase, n_drivers: int, device, group=None, src_rank: int = 0, warmup: int = 2,
                 overlap: bool = True):
        self.sh = ShardedGbase(gbase, group=group, src_rank=src_rank)
        self.G = gbase
        dev = torch.device(device)
        self.xs = torch.zeros((1, 3, 512, 512), dtype=torch.float32, device=dev)
        self.xd = torch.zeros((n_drivers, 3, 512, 512), dtype=torch.float32, device=dev)
        self.flat = torch.zeros(STATE_NUMEL, dtype=torch.float32, device=dev)
        self.is_src = self.sh.rank == src_rank
        cur = torch.cuda.current_stream(dev)
        side = torch.cuda.Stream(dev)
        side.wait_stream(cur)
        with torch.cuda.stream(side), torch.no_grad():
            for _ in range(warmup):     # plans packed, TMA encoder resolved, cuBLAS / allocator warm
                src = gbase.encode_source(self.xs)
                pack_source(src, self.flat)
                gbase.drive_render(unpack_source(self.flat), gbase.drive_motion(self.xd))
        cur.wait_stream(side)
        torch.cuda.synchronize(dev)
        l0 = ops.LAUNCHES
        self.g_pre = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_pre), torch.no_grad():
            main = torch.cuda.current_stream(dev)
            if self.is_src and overlap:
                fork = torch.cuda.Stream(dev)
                fork.wait_stream(main)
                with torch.cuda.stream(fork):
                    pack_source(gbase.encode_source(self.xs), self.flat)
                self.motion = gbase.drive_motion(self.xd)
                main.wait_stream(fork)
            else:
                if self.is_src:
                    pack_source(gbase.encode_source(self.xs), self.flat)
                self.motion = gbase.drive_motion(self.xd)
        self.pre_launches = ops.LAUNCHES - l0
        l0 = ops.LAUNCHES
        self.g_render = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_render), torch.no_grad():
            self.out = gbase.drive_render(unpack_source(self.flat), self.motion)
        self.render_launches = ops.LAUNCHES - l0
        # motion encoder alone (further chunks of driver frames that share the already encoded source): results are
        # copied into the static tensors that the render graph reads
        l0 = ops.LAUNCHES
        self.g_motion = torch.cuda.CUDAGraph()
        with torch.cuda.graph(self.g_motion), torch.no_grad():
            for dst, src in zip(self.motion, gbase.drive_motion(self.xd)):
                dst.copy_(src)
        self.motion_launches = ops.LAUNCHES - l0

    @torch.no_grad()
    def step(self, xs: torch.Tensor, xd_local: torch.Tensor):
        """Returns (xhat, pyramids): STATIC tensors, valid until the next call."""
        self.xd.copy_(xd_local, non_blocking=True)
        if self.is_src:
            self.xs.copy_(xs, non_blocking=True)
        self.g_pre.replay()
        ops._count(self.pre_launches)
        if self.sh.world > 1:
            self.sh.dist.broadcast(self.flat, self.sh.src_rank, group=self.sh.group)
        self.g_render.replay()
        ops._count(self.render_launches)
        return self.out

    @torch.no_grad()
    def step_chunks(self, xs: torch.Tensor, xd_local: torch.Tensor, sink=None):
        """One source, `xd_local.shape[0]` = k * n_drivers driver frames on this rank (BASELINE config 3 with 256 / N
        frames per GPU): the source is encoded and broadcast ONCE, then every chunk of n_drivers frames runs the motion
        encoder and the render graph.  `sink(i, out)` is called with the static outputs of chunk i before the next chunk
        overwrites them (None: outputs are dropped, as in a throughput run)."""
        n = self.xd.shape[0]
        if xd_local.shape[0] % n:
            raise RuntimeError(f"step_chunks: {xd_local.shape[0]} driver frames are not a multiple of the chunk size {n}")
        for i in range(xd_local.shape[0] // n):
            chunk = xd_local[i * n:(i + 1) * n]
            if i == 0:
                out = self.step(xs, chunk)
            else:
                self.xd.copy_(chunk, non_blocking=True)
                self.g_motion.replay()
                self.g_render.replay()
                ops._count(self.motion_launches + self.render_launches)
                out = self.out
            if sink is not None:
                sink(i, out)
        return out


# ----------------------------------------------------------------------------------------------------- training (row f-2)
class GradBucket:
    """All gradients of a module as views into ONE flat fp32 buffer, so data-parallel training needs no gather / scatter copies
    (BASELINE config 5: `train_base` on 8 GPUs, data-parallel; the reference itself is single-process, train.py:129-356).

    `attach()` points every `param.grad` at its slice of the buffer (autograd accumulates into an existing `.grad` in place, so the
    backward pass fills the buffer); `zero()` clears it with one memset; `all_reduce_mean()` averages it over the process group.
    Use `optimizer.zero_grad(set_to_none=False)` or just `bucket.zero()`: setting grads to None would detach the views.

    `n_buckets` > 1 overlaps the collective with the backward pass: the buffer is cut into contiguous ranges of about equal size
    (parameters in registration order, which the backward pass walks roughly in reverse), and the moment the last gradient of a
    range has been accumulated (`register_post_accumulate_grad_hook`) its all-reduce is enqueued on a communication stream that
    runs beside the rest of the backward pass; `all_reduce_mean()` then only reduces what is left, joins the stream and scales.
    Under CUDA-graph capture the communication stream becomes a parallel branch of the graph."""

    def __init__(self, module: torch.nn.Module, group=None, n_buckets: int = 1):
        import torch.distributed as dist
        self.params = [p for p in module.parameters() if p.requires_grad]
        if not self.params:
            raise RuntimeError("GradBucket: the module has no trainable parameters")
        dev = self.params[0].device
        sizes = [p.numel() for p in self.params]
        # 16-byte aligned slices: vectorised optimizer kernels and NCCL's 16-byte chunks stay aligned
        self.offsets, total = [], 0
        for n in sizes:
            self.offsets.append(total)
            total += (n + 3) // 4 * 4
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.dist = dist if (dist.is_available() and dist.is_initialized()) else None
        self.group = group
        self.world = self.dist.get_world_size(group) if self.dist else 1
        # contiguous ranges of ~equal size: (lo, hi, number of parameters inside)
        n_buckets = max(1, min(int(n_buckets), len(self.params)))
        self.ranges, self._bucket_of = [], {}
        target, lo, count = total / n_buckets, 0, 0
        for i, (p, off) in enumerate(zip(self.params, self.offsets)):
            end = self.offsets[i + 1] if i + 1 < len(self.params) else total
            self._bucket_of[id(p)] = len(self.ranges)
            count += 1
            if (end >= target * (len(self.ranges) + 1) and len(self.ranges) < n_buckets - 1) or i + 1 == len(self.params):
                self.ranges.append((lo, end, count))
                lo, count = end, 0
        self._pending = [c for _, _, c in self.ranges]
        self._launched = [False] * len(self.ranges)
        self._comm = torch.cuda.Stream(dev) if (dev.type == "cuda" and len(self.ranges) > 1) else None
        self._hooks = []
        if len(self.ranges) > 1:
            for p in self.params:
                self._hooks.append(p.register_post_accumulate_grad_hook(self._on_grad))
        self.attach()

    def attach(self) -> None:
        for p, off in zip(self.params, self.offsets):
            if p.dtype != torch.float32 or p.device != self.flat.device:
                raise RuntimeError("GradBucket: fp32 parameters on one device expected")
            p.grad = self.flat[off:off + p.numel()].view_as(p)

    def zero(self) -> None:
        self.flat.zero_()
        self._pending = [c for _, _, c in self.ranges]
        self._launched = [False] * len(self.ranges)

    def _on_grad(self, p) -> None:
        b = self._bucket_of[id(p)]
        self._pending[b] -= 1
        if self._pending[b] == 0:
            self._launch(b)

    def _launch(self, b: int) -> None:
        if self._launched[b]:
            return
        self._launched[b] = True
        if self.dist is None or self.world == 1:
            return
        lo, hi, _ = self.ranges[b]
        if self._comm is None:
            self.dist.all_reduce(self.flat[lo:hi], op=self.dist.ReduceOp.SUM, group=self.group)
            return
        self._comm.wait_stream(torch.cuda.current_stream(self.flat.device))      # the range's gradients are complete here
        with torch.cuda.stream(self._comm):
            self.dist.all_reduce(self.flat[lo:hi], op=self.dist.ReduceOp.SUM, group=self.group)

    def all_reduce_mean(self) -> None:
        """Sum over ranks, then 1 / world: ONE collective (n_buckets = 1), or whatever the backward pass has not already sent
        (ranges holding a parameter that received no gradient, e.g. the unused `adaptive_matrix_beta`).  No-op in one process."""
        if self.dist is None or self.world == 1:
            return
        for b in range(len(self.ranges)):
            self._launch(b)
        if self._comm is not None:
            torch.cuda.current_stream(self.flat.device).wait_stream(self._comm)
        self.flat.mul_(1.0 / self.world)


class DataParallelTrainer:
    """One process per GPU: every rank runs `Gbase.train()` forward + backward on its own (source, driver) pairs, the gradients
    meet in ONE all-reduce (`GradBucket`), every rank applies the same optimizer step (weights stay bit-identical across ranks as
    long as they start equal).  BatchNorm statistics stay per-rank, as `DistributedDataParallel` does by default.
    `loss_fn(pred, pyramids, xs, xd) -> scalar` stands in for the reference's out-of-scope losses.

    `graph=True`: after `warmup` eager iterations the whole iteration -- forward, backward, the all-reduce and the optimizer step --
    is captured into ONE CUDA graph per input shape and replayed (at batch 1 the eager iteration is bound by the host enqueueing
    ~7 000 launches: 160-250 ms eager vs 77 ms replayed on one B200).  The optimizer must be capturable
    (`torch.optim.AdamW(..., capturable=True, fused=True)`: the fused variant also saves ~15 ms per iteration over the foreach one), `loss_fn` must not synchronise, inputs are copied into static buffers."""

    def __init__(self, gbase, optimizer_factory, loss_fn=None, group=None, graph: bool = False, warmup: int = 2,
                 n_buckets: int = 1):
        self.G = gbase.train()
        self.bucket = GradBucket(gbase, group, n_buckets)        # n_buckets > 1: all-reduce overlapped with the backward pass
        self.opt = optimizer_factory(self.bucket.params)
        self.loss_fn = loss_fn or (lambda pred, pyr, xs, xd: (pred - xd).abs().mean())
        self.graph, self.warmup = bool(graph), int(warmup)
        self._graphs: Dict[tuple, tuple] = {}
        self._seen: Dict[tuple, int] = {}

    def _iteration(self, xs: torch.Tensor, xd: torch.Tensor) -> torch.Tensor:
        self.bucket.zero()
        with torch.enable_grad():
            pred, pyr = self.G(xs, xd)
            loss = self.loss_fn(pred, pyr, xs, xd)
            loss.backward()
        self.bucket.all_reduce_mean()
        self.opt.step()
        return loss.detach()

    def step(self, xs: torch.Tensor, xd: torch.Tensor) -> torch.Tensor:
        if not self.graph:
            return self._iteration(xs, xd)
        key = (tuple(xs.shape), tuple(xd.shape), str(xs.device))
        ent = self._graphs.get(key)
        if ent is None:
            n = self._seen.get(key, 0)
            self._seen[key] = n + 1
            if n < self.warmup:                       # eager iterations first: lazy initialisations, optimizer state, NCCL
                return self._iteration(xs, xd)
            sx, sd_ = xs.clone(), xd.clone()
            side = torch.cuda.Stream(xs.device)
            side.wait_stream(torch.cuda.current_stream(xs.device))
            with torch.cuda.stream(side):             # THIS call's iteration, eagerly on a side stream (torch's capture recipe)
                loss_now = self._iteration(sx, sd_)
            torch.cuda.current_stream(xs.device).wait_stream(side)
            torch.cuda.synchronize(xs.device)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):                 # recorded, not executed: the following calls replay it
                loss = self._iteration(sx, sd_)
            self._graphs[key] = (g, sx, sd_, loss)
            return loss_now
        g, sx, sd_, loss = ent
        sx.copy_(xs)
        sd_.copy_(xd)
        g.replay()
        return loss.clone()
