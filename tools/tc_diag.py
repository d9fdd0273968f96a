"""Diagnostics for mp_conv_tc on a real B200 (developer tool, run via gpurun): each case runs in-process but
prints enough structure (error by tile row, by 16-column chunk, delta probes) to locate descriptor / swizzle bugs."""
import math
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from megaportrait_hack_b200 import lib, ops  # noqa: E402

lib.build()
DEV = "cuda"


def run(name, N, Cin, Cout, D, H, W, k, probe=False):
    g = torch.Generator().manual_seed(5)
    fan = Cin * k[0] * k[1] * k[2]
    x = torch.randn(N, Cin, D, H, W, generator=g)
    w = torch.randn(Cout, Cin, *k, generator=g) / math.sqrt(fan)
    b = torch.randn(Cout, generator=g) * 0.1
    if probe:   # delta probe: identity 1x1 => output must equal input
        x.zero_(); w.zero_(); b.zero_()
        for c in range(min(Cin, Cout)):
            w[c, c, k[0] // 2, k[1] // 2, k[2] // 2] = 1.0
        x[0, 3 % Cin, D // 2, H // 3, W // 5] = 1.0
        x[0, (Cin - 1), 0, 0, 0] = 2.0
        x[N - 1, 17 % Cin, D - 1, H - 1, W - 1] = 3.0
    ref = F.conv3d(x, w, b, padding=tuple(i // 2 for i in k))
    a = ops.from_nchw(x.to(DEV), f32=False, split=True)
    pw = ops.pack_conv(w, b, DEV)
    out = {}
    for mode in ("simt", "tc"):
        try:
            o, _ = ops.conv(a, pw, f32=True, mode=mode)
            torch.cuda.synchronize()
            out[mode] = o.f32.permute(0, 4, 1, 2, 3).contiguous().cpu()
        except Exception as e:  # noqa: BLE001
            print(f"[{name}] {mode}: EXCEPTION {e}")
            return False
    scale = ref.abs().max().item()
    ok = True
    for mode, got in out.items():
        err = (got - ref).abs()
        rel = err.max().item() / scale
        bad = (err > 1e-4 * scale).float().mean().item()
        status = "ok" if rel < 3e-5 else "MISMATCH"
        print(f"[{name}] {mode}: rel={rel:.3e} bad_frac={bad:.4f} nan={torch.isnan(got).any().item()} "
              f"absmax_got={got.abs().max().item():.4f} absmax_ref={scale:.4f} {status}")
        if status != "ok":
            ok = False
            if mode == "tc":
                e = err.reshape(N, Cout, -1)
                per_c = e.amax(dim=(0, 2))
                chunks = per_c.reshape(-1, min(16, Cout)).amax(1) if Cout % 16 == 0 else per_c
                print("   err by 16-channel chunk:", [f"{v:.2e}" for v in chunks.tolist()][:32])
                pos = e.amax(dim=(0, 1))
                print("   err by position%8:", [f"{pos[i::8].max().item():.2e}" for i in range(8)])
                print("   err by position//(S/8):", [f"{v.max().item():.2e}" for v in pos.chunk(8)])
                if probe:
                    nz = (got.abs() > 1e-6).nonzero()[:12].tolist()
                    rz = (ref.abs() > 1e-6).nonzero()[:12].tolist()
                    print("   got nonzeros:", [(i, round(got[tuple(i)].item(), 3)) for i in nz])
                    print("   ref nonzeros:", [(i, round(ref[tuple(i)].item(), 3)) for i in rz])
    return ok


CASES = [
    ("probe_1x1_c64", 1, 64, 64, 1, 16, 64, (1, 1, 1), True),
    ("probe_3x3_c64", 1, 64, 64, 1, 16, 64, (1, 3, 3), True),
    ("rand_1x1_c64", 1, 64, 64, 1, 16, 64, (1, 1, 1), False),
    ("rand_3x3_c64", 1, 64, 64, 1, 16, 64, (1, 3, 3), False),
    ("rand_3x3_c128_n128", 2, 128, 128, 1, 32, 32, (1, 3, 3), False),
    ("probe_1x1_c32", 1, 32, 32, 1, 16, 64, (1, 1, 1), True),
    ("rand_3x3_c96", 1, 96, 96, 1, 16, 64, (1, 3, 3), False),
    ("rand_3x3x3_c96", 1, 96, 96, 4, 16, 32, (3, 3, 3), False),
    ("probe_1x1_c16", 1, 16, 16, 1, 16, 64, (1, 1, 1), True),
    ("rand_3x3_c16", 1, 16, 32, 1, 16, 64, (1, 3, 3), False),
    ("rand_3x3_c512", 1, 512, 512, 1, 64, 64, (1, 3, 3), False),
]

if __name__ == "__main__":
    assert lib.load().mp_device_supported() == 1
    results = {}
    for c in CASES:
        try:
            results[c[0]] = run(*c)
        except Exception as e:  # noqa: BLE001
            print(f"[{c[0]}] fatal: {e}")
            results[c[0]] = False
            break
    print("SUMMARY", results)
