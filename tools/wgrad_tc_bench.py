"""Weight-gradient kernels timed alone (CUDA events, L2 flushed between repetitions): tcgen05 (`mp_conv_wgrad_tc`) vs the
first-generation mma.sync kernel.  Usage (GPU box): python tools/wgrad_tc_bench.py [out.json]"""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import __graft_entry__ as entry  # noqa: E402


def main():
    entry.build()
    from megaportrait_hack_b200 import ops
    dev = "cuda"
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    shapes = [("96->96 3x3x3 @16x64x64", 1, 96, 96, (16, 64, 64), (3, 3, 3)),
              ("512->512 3x3 @64^2 x8", 8, 512, 512, (1, 64, 64), (1, 3, 3)),
              ("64->64 3x3 @512^2", 1, 64, 64, (1, 512, 512), (1, 3, 3)),
              ("128->128 3x3 @512^2", 1, 128, 128, (1, 512, 512), (1, 3, 3)),
              ("256->256 3x3 @256^2", 1, 256, 256, (1, 256, 256), (1, 3, 3)),
              ("768->768 3x3x3 @2x8x8", 1, 768, 768, (2, 8, 8), (3, 3, 3)),
              ("16->64 7x7 @512^2", 1, 16, 64, (1, 512, 512), (1, 7, 7))]
    rows = []
    for name, N, Ci, Co, sp, k in shapes:
        x = ops._to_cl_act(torch.randn(N, Ci, *sp, device=dev))
        g = ops._to_cl_act(torch.randn(N, Co, *sp, device=dev))
        ops.ensure_split(x), ops.ensure_split(g)
        fl = 2 * N * sp[0] * sp[1] * sp[2] * Ci * Co * k[0] * k[1] * k[2]
        row = {"shape": name, "gflop": fl / 1e9}
        for mode in ("tc", "mma"):
            ops.WGRAD_TC = mode == "tc"
            for _ in range(2):
                ops.conv_weight_grad(x, g, k)
            ts = []
            for _ in range(5):
                flush.zero_()
                t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                t0.record()
                ops.conv_weight_grad(x, g, k)
                t1.record()
                torch.cuda.synchronize()
                ts.append(t0.elapsed_time(t1))
            ms = sorted(ts)[len(ts) // 2]
            row[mode + "_ms"] = ms
            row[mode + "_useful_tflops"] = fl / (ms * 1e-3) / 1e12
        rows.append(row)
        print(json.dumps(row))
    if len(sys.argv) > 1:
        json.dump(rows, open(sys.argv[1], "w"), indent=1)


if __name__ == "__main__":
    main()
