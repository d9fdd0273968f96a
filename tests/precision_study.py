"""Operand-format study behind the precision policy of DESIGN.md section 4 (NOT a test: pytest does not collect it).

    python tests/precision_study.py [formats] [stages] [emtn] [fp8]

It runs the CPU oracle (test infrastructure, `oracle/gbase_oracle.py`) with the operands of every convolution rounded
the way a candidate tensor-core format would round them, and prints the end-to-end RGB error and the per-stage errors.
Results on the seeded weights / inputs (torch 2.11 CPU; RGB max-abs, budget 1e-3):

  formats  split-bf16 both operands (3 passes, the default)          1.8e-5
           fp16 activations, exact weights (2 passes, everywhere)     1.2e-3     <- over budget
           fp16 both (1 pass)                                         1.5e-3
           tf32 truncation both (hardware TF32)                       5.7e-3
           bf16 both (1 pass)                                         1.4e-2
  stages   fp16 activations in ONE stage only:  G2d 1.0e-3, Eapp 3.8e-4, G3d 3.6-6.7e-4, FlowField 2-3e-4, Emtn 9e-6
  emtn     Emtn with every activation STORED as fp16 + fp16 hi/scaled-lo weights (what MP_PREC_F16X2 does):
           RGB 2.0e-5; R/t/z 6e-6..1.3e-5 relative L2; w_c2d 1.9e-5   -> adopted (the trunks end in global average pools)
           same format for the ResNet-50 descriptor: es 1.8e-4, w_s2c 1.6e-4 relative L2 -> NOT adopted
  fp8      fp16 main product + both cross terms in FP8 (2 pass-units):  e4m3 5.3-6.2e-5, e5m2 1.3-1.7e-4 -> future work
"""
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)
import gbase_oracle as O  # noqa: E402
from megaportrait_hack_b200 import seeded  # noqa: E402

CONV2D, CONV3D, RELU, BN = F.conv2d, F.conv3d, F.relu, O._bn


def f16(x):
    return x.half().float()


def bf16(x):
    return x.bfloat16().float()


def split(x, cast, scale=1.0):
    hi = cast(x)
    return hi + cast((x - hi) * scale) / scale


def tf32_trunc(x):
    return (x.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)


def inputs(seed=1, n=1):
    g = torch.Generator().manual_seed(seed)
    return torch.rand(1, 3, 512, 512, generator=g), torch.rand(n, 3, 512, 512, generator=g)


class patched:
    """Round conv operands with (fa, fw) while `active()` is true."""

    def __init__(self, fa, fw, active=lambda: True, store=None):
        self.fa, self.fw, self.active, self.store = fa, fw, active, store

    def __enter__(self):
        def mk(fn):
            def conv(x, w, b=None, *a, **k):
                if self.active():
                    x, w = self.fa(x.contiguous()), self.fw(w)
                return fn(x, w, b, *a, **k)
            return conv
        F.conv2d, F.conv3d = mk(CONV2D), mk(CONV3D)
        if self.store is not None:      # activations are STORED rounded (after ReLU and after every folded BatchNorm)
            F.relu = lambda x, *a, **k: (self.store(RELU(x, *a, **k)) if self.active() else RELU(x, *a, **k))
            O._bn = lambda x, sd, p: (self.store(BN(x, sd, p)) if self.active() else BN(x, sd, p))
        return self

    def __exit__(self, *exc):
        F.conv2d, F.conv3d, F.relu, O._bn = CONV2D, CONV3D, RELU, BN


def scoped(names):
    """Context flag that is true only inside the named oracle stage functions."""
    depth = {"n": 0}
    saved = {}

    def wrap(fn):
        def f(*a, **k):
            depth["n"] += 1
            try:
                return fn(*a, **k)
            finally:
                depth["n"] -= 1
        return f
    for n in names:
        saved[n] = getattr(O, n)
        setattr(O, n, wrap(saved[n]))
    return (lambda: depth["n"] > 0), (lambda: [setattr(O, n, f) for n, f in saved.items()])


def report(tag, rgb, ref, st=None, st0=None, keys=()):
    print(f"{tag:58s} RGB max-abs {(rgb - ref).abs().max().item():.3e}", flush=True)
    for k in keys:
        a, b = st[k].float(), st0[k].float()
        print(f"    {k:10s} rel-L2 {((a - b).norm() / b.norm()).item():.3e}   max-abs/abs-max "
              f"{((a - b).abs().max() / b.abs().max()).item():.3e}")


def main():
    what = set(sys.argv[1:]) or {"formats", "stages", "emtn", "fp8"}
    torch.set_num_threads(os.cpu_count() or 8)
    sd = seeded.seeded_state_dict(seed=0)
    ident = lambda x: x
    with torch.no_grad():
        if "formats" in what:
            xs, xd = inputs()
            ref, _ = O.gbase_forward(xs, xd, sd)
            for tag, fa, fw in (("split-bf16 both (3 passes)", lambda x: split(x, bf16), lambda w: split(w, bf16)),
                                ("fp16 activations, exact weights (2 passes)", f16, ident),
                                ("exact activations, fp16 weights (2 passes)", ident, f16),
                                ("fp16 both (1 pass)", f16, f16),
                                ("tf32 truncation both", tf32_trunc, tf32_trunc),
                                ("bf16 both (1 pass)", bf16, bf16)):
                with patched(fa, fw):
                    rgb, _ = O.gbase_forward(xs, xd, sd)
                report(tag, rgb, ref)
        if "stages" in what:
            xs, xd = inputs()
            ref, _ = O.gbase_forward(xs, xd, sd)
            for stage in ("g2d", "eapp", "g3d", "flowfield", "emtn"):
                active, restore = scoped([stage])
                with patched(f16, ident, active):
                    rgb, _ = O.gbase_forward(xs, xd, sd)
                restore()
                report(f"fp16 activations inside {stage} only", rgb, ref)
        if "emtn" in what:
            xs, xd = inputs(n=2)
            ref, _, st0 = O.gbase_forward_shared_source(xs, xd, sd, stages=True)
            keys = ("es", "Rs", "ts", "zs", "Rd", "td", "zd", "w_s2c", "w_c2d", "vc2d", "projected")
            for stages_ in (["emtn"], ["custom_resnet50"]):
                active, restore = scoped(stages_)
                with patched(f16, lambda w: split(w, f16, 2048.0), active, store=f16):
                    rgb, _, st = O.gbase_forward_shared_source(xs, xd, sd, stages=True)
                restore()
                report(f"MP_PREC_F16X2 storage format inside {stages_[0]}", rgb, ref, st, st0, keys)
        if "fp8" in what:
            for dt, lim, tag in ((torch.float8_e4m3fn, 448.0, "e4m3"), (torch.float8_e5m2, 57344.0, "e5m2")):
                q = lambda x: x.clamp(-lim, lim).to(dt).float()

                def mk(fn):
                    def conv(x, w, b=None, *a, **k):
                        x = x.contiguous()
                        xh, wh = f16(x), f16(w)
                        sw = 2.0 ** (-math.floor(math.log2(max(w.abs().max().item(), 1e-30))))
                        corr = fn(q(x), q((w - wh) * 2048 * sw) / sw, None, *a, **k) + \
                            fn(q((x - xh) * 2048), q(w * sw) / sw, None, *a, **k)
                        return fn(xh, wh, b, *a, **k) + corr / 2048
                    return conv
                for seed in (1, 7):
                    xs, xd = inputs(seed)
                    ref, _ = O.gbase_forward(xs, xd, sd)
                    F.conv2d, F.conv3d = mk(CONV2D), mk(CONV3D)
                    try:
                        rgb, _ = O.gbase_forward(xs, xd, sd)
                    finally:
                        F.conv2d, F.conv3d = CONV2D, CONV3D
                    report(f"fp16 x fp16 + two {tag} cross terms (2 pass-units), seed {seed}", rgb, ref)


if __name__ == "__main__":
    main()
