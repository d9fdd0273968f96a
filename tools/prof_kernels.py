"""Stand-alone launches of the hot kernels at bench shapes (developer tool for ncu / CUDA-event timing via gpurun).

    python tools/prof_kernels.py [--reps 5]     # prints per-kernel event timings, TFLOP/s / GB/s
    ncu --set full -k regex:k_conv_tc -c 3 python tools/prof_kernels.py --reps 1
"""
import argparse
import math
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from megaportrait_hack_b200 import lib, ops  # noqa: E402

lib.build()
DEV = "cuda"
FLUSH = None


def flush_l2():
    global FLUSH
    if FLUSH is None:
        FLUSH = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=DEV)
    FLUSH.fill_(1.0)


def timeit(fn, reps):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        flush_l2()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return min(ts), sum(ts) / len(ts)


def conv_case(name, N, Cin, Cout, D, H, W, k, reps, mode="tc", act=ops.ACT_RELU, split_out=True):
    g = torch.Generator().manual_seed(0)
    a = ops.Act((N, D, H, W, Cin), f32=torch.randn(N, D, H, W, Cin, generator=g).to(DEV))
    ops.ensure_split(a)
    a.f32 = None
    w = torch.randn(Cout, Cin, *k, generator=g) / math.sqrt(Cin * k[0] * k[1] * k[2])
    pw = ops.pack_conv(w, torch.zeros(Cout), DEV)
    fn = lambda: ops.conv(a, pw, act=act, f32=not split_out, split=split_out, mode=mode)
    best, avg = timeit(fn, reps)
    fl = 2.0 * N * D * H * W * Cout * Cin * k[0] * k[1] * k[2]
    print(f"{name:28s} {mode}: best {best:8.3f} ms avg {avg:8.3f} ms  useful {fl / best / 1e9:8.1f} TFLOP/s "
          f"(raw bf16 x3 {3 * fl / best / 1e9:8.1f})", flush=True)


def conv_case_h(name, N, Cin, Cout, H, W, k, reps, stride=1, res=False):
    """MP_PREC_F16X2 convolution (fp16 activation plane in / out) at a motion-encoder shape."""
    g = torch.Generator().manual_seed(0)
    a = ops.Act((N, 1, H, W, Cin), h16=torch.randn(N, 1, H, W, Cin, generator=g).to(DEV).half())
    w = torch.randn(Cout, Cin, k, k, generator=g) / math.sqrt(Cin * k * k)
    pw = ops.pack_conv(w, torch.zeros(Cout), DEV, prec=ops.PREC_F16X2)
    r = ops.Act((N, 1, H // stride, W // stride, Cout),
                h16=torch.randn(N, 1, H // stride, W // stride, Cout, generator=g).to(DEV).half()) if res else None
    fn = lambda: ops.conv(a, pw, res=r, act=ops.ACT_RELU, f32=False, h16=True, stride=stride, mode="tc")
    best, avg = timeit(fn, reps)
    fl = 2.0 * N * (H // stride) * (W // stride) * Cout * Cin * k * k
    by = N * H * W * Cin * 2 + N * (H // stride) * (W // stride) * Cout * 2 * (2 if res else 1)
    print(f"{name:28s} f16x2: best {best:8.3f} ms avg {avg:8.3f} ms  useful {fl / best / 1e9:8.1f} TFLOP/s "
          f"(raw x2 {2 * fl / best / 1e9:8.1f})  activations {by / best / 1e6:8.1f} GB/s", flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--only", default="")
    args = ap.parse_args()
    r = args.reps
    sel = lambda n: (not args.only) or args.only in n
    if sel("conv"):
        conv_case("g2d_512_64x64_b32", 32, 512, 512, 1, 64, 64, (1, 3, 3), r)
        conv_case("g2d_256_128x128_b32", 32, 256, 256, 1, 128, 128, (1, 3, 3), r)
        conv_case("g2d_128to64_512x512_b8", 8, 128, 64, 1, 512, 512, (1, 3, 3), r)
        conv_case("g2d_in_1x1_96to512_b32", 32, 96, 512, 1, 64, 64, (1, 1, 1), r)
        conv_case("g2d_64_512x512_b32", 32, 64, 64, 1, 512, 512, (1, 3, 3), r, split_out=False)
        conv_case("g2d_sc_128to64_1x1_b32", 32, 128, 64, 1, 512, 512, (1, 1, 1), r, split_out=False, act=ops.ACT_NONE)
        conv_case("g2d_128to64_512x512_b32", 32, 128, 64, 1, 512, 512, (1, 3, 3), r)
        conv_case("g2d_up2_256to128_b32", 32, 256, 128, 1, 256, 256, (1, 3, 3), r)
        conv_case("g2d_head_64to3_b32", 32, 64, 3, 1, 512, 512, (1, 3, 3), r, split_out=False, act=ops.ACT_SIGMOID)
        conv_case("eapp_64to128_512_b1", 1, 64, 128, 1, 512, 512, (1, 3, 3), r, split_out=False, act=ops.ACT_NONE)
        conv_case("g2d_up1_512to256_b32", 32, 512, 256, 1, 128, 128, (1, 3, 3), r)
        conv_case("vol96_16x64x64_b1", 1, 96, 96, 16, 64, 64, (3, 3, 3), r, split_out=False, act=ops.ACT_NONE)
        conv_case("eapp_128_512x512_b1", 1, 128, 128, 1, 512, 512, (1, 3, 3), r, split_out=False, act=ops.ACT_NONE)
        conv_case("g3d_768_2x8x8_b1", 1, 768, 768, 2, 8, 8, (3, 3, 3), r, split_out=False, act=ops.ACT_NONE)
    if sel("small"):
        conv_case("g2d_128to64_512x512_b32", 32, 128, 64, 1, 512, 512, (1, 3, 3), r)
        conv_case("g2d_64_512x512_b32", 32, 64, 64, 1, 512, 512, (1, 3, 3), r, split_out=False)
        conv_case("g2d_up2_256to128_b32", 32, 256, 128, 1, 256, 256, (1, 3, 3), r)
        conv_case("r18_64_256x256_b32", 32, 64, 64, 1, 256, 256, (1, 3, 3), r)
        conv_case("vol96_16x64x64_b1", 1, 96, 96, 16, 64, 64, (3, 3, 3), r, split_out=False, act=ops.ACT_NONE)
    if sel("f16"):
        conv_case_h("emtn_stem_k32_512_b32", 32, 32, 128, 512, 512, 1, r)
        conv_case_h("r18_64_256x256_b32_h", 32, 64, 64, 256, 256, 3, r)
        conv_case_h("r18_64_256x256_b32_h_res", 32, 64, 64, 256, 256, 3, r, res=True)
        conv_case_h("r18_128_128x128_b32_h", 32, 128, 128, 128, 128, 3, r)
        conv_case_h("r18_512_32x32_b32_h", 32, 512, 512, 32, 32, 3, r)
    if sel("up"):
        g = torch.Generator().manual_seed(3)
        for (C, H) in ((512, 64), (256, 128), (128, 256)):
            a = ops.Act((32, 1, H, H, C), f32=torch.randn(32, 1, H, H, C, generator=g).to(DEV))
            ops.ensure_split(a)
            a.f32 = None
            best, avg = timeit(lambda: ops.upsample2x_linear(a, 1, f32=False, split=True), r)
            by = 32 * H * H * C * 4 * 5
            print(f"upsample2x_split_{C}x{H}       best {best:8.3f} ms avg {avg:8.3f} ms  {by / best / 1e6:8.1f} GB/s", flush=True)
    if sel("head"):
        g = torch.Generator().manual_seed(4)
        a = ops.Act((32, 1, 512, 512, 64), f32=torch.randn(32, 1, 512, 512, 64, generator=g).to(DEV))
        ab = torch.rand(32, 64, 2, generator=g).to(DEV)
        w, b = torch.randn(3, 64, 3, 3, generator=g) / 24, torch.zeros(3)
        best, avg = timeit(lambda: ops.gn_relu_conv3x3_head(a, ab, w, b), r)
        by = 32 * 512 * 512 * (64 + 3) * 4
        print(f"gn_relu_conv3x3_head_b32     best {best:8.3f} ms avg {avg:8.3f} ms  {by / best / 1e6:8.1f} GB/s", flush=True)
    if sel("q8"):
        # G2d identity res-block convolution in the fp16 + FP8 cross-term mode
        g = torch.Generator().manual_seed(5)
        xcl = torch.randn(32, 1, 64, 64, 512, generator=g)
        a = ops.Act(tuple(xcl.shape), h16=xcl.half().to(DEV), q8=ops.q8_planes(xcl).to(DEV))
        w = torch.randn(512, 512, 3, 3, generator=g) / math.sqrt(512 * 9)
        pw = ops.pack_conv(w, torch.zeros(512), DEV, prec=ops.PREC_F16_Q8)
        best, avg = timeit(lambda: ops.conv(a, pw, res=a, act=ops.ACT_RELU, f32=False, hq=True), r)
        fl = 2.0 * 32 * 64 * 64 * 512 * 512 * 9
        print(f"g2d_512_64x64_b32_q8         best {best:8.3f} ms avg {avg:8.3f} ms  useful {fl / best / 1e9:8.1f} TFLOP/s "
              f"(pass-units x2 {2 * fl / best / 1e9:8.1f})", flush=True)
    if sel("simt"):
        conv_case("g2d_512_64x64_b4", 4, 512, 512, 1, 64, 64, (1, 3, 3), r, mode="simt")
    if sel("warp"):
        g = torch.Generator().manual_seed(1)
        for N, Nv, sum_d, tag in ((32, 1, True, "warp_sumD_b32_shared"), (1, 1, False, "warp_b1"),
                                  (32, 32, False, "warp_b32_own")):
            v = ops.Act((Nv, 16, 64, 64, 96), f32=torch.randn(Nv, 16, 64, 64, 96, generator=g).to(DEV))
            for label, scale in (("faithful", 1.0), ("spread", 40.0)):
                em = (torch.rand(N, 16, 16, 16, 3, generator=g) * scale).to(DEV)
                th = torch.eye(4)[:3][None].repeat(N, 1, 1).contiguous().to(DEV) * (1.0 if scale == 1.0 else 20.0)
                fn = lambda: ops.warp_fused(v, em, th, sum_d=sum_d, f32=True, split=False)
                best, avg = timeit(fn, r)
                by = Nv * 16 * 64 * 64 * 96 * 4 + N * 16 ** 3 * 3 * 4 + N * (1 if sum_d else 16) * 64 * 64 * 96 * 4
                print(f"{tag + '_' + label:28s} best {best:8.3f} ms avg {avg:8.3f} ms  algorithmic {by / best / 1e6:8.1f} GB/s",
                      flush=True)
    if sel("grid"):
        import torch.nn.functional as F
        g = torch.Generator().manual_seed(2)
        for B in (1, 8):
            v = torch.randn(B, 96, 16, 64, 64, generator=g).to(DEV)
            zz, yy, xx = torch.meshgrid(torch.linspace(-1, 1, 16), torch.linspace(-1, 1, 64),
                                        torch.linspace(-1, 1, 64), indexing="ij")
            base = torch.stack((xx, yy, zz), -1)[None].repeat(B, 1, 1, 1, 1)
            grids = {"spread": base + (torch.rand(B, 16, 64, 64, 3, generator=g) - 0.5) * 0.2,
                     "adversarial": (torch.rand(B, 16, 64, 64, 3, generator=g) - 0.5) * 3.0}
            for label, grid in grids.items():
                gd = grid.to(DEV)
                by = B * 51118080
                best, avg = timeit(lambda: ops.grid_sample3d(v, gd), r)
                bd, _ = timeit(lambda: ops.grid_sample3d(v, gd, direct=True), r)
                tb, _ = timeit(lambda: F.grid_sample(v, gd, mode="bilinear", padding_mode="border", align_corners=True), r)
                print(f"grid_sample3d_b{B}_{label:12s} best {best:8.3f} ms  algorithmic {by / best / 1e6:8.1f} GB/s   "
                      f"(direct NCDHW kernel {bd:8.3f} ms, {by / bd / 1e6:8.1f} GB/s; ATen kernel: {tb:8.3f} ms, "
                      f"{by / tb / 1e6:8.1f} GB/s)", flush=True)


if __name__ == "__main__":
    main()
