"""TEST INFRASTRUCTURE ONLY -- golden vectors for `Genh` from the REAL reference class.

    python oracle/make_golden_genh.py

The reference's `Genh()` raises TypeError as written (`ResBlock2D(64)`, model.py:1354, against `ResBlock2D.__init__(self,
in_channels, out_channels, downsample=False)`, model.py:601).  The minimal repair -- the one DESIGN.md documents -- is applied
while the reference class is constructed: `ResBlock2D(c)` means `ResBlock2D(c, c)`.  Everything else (layer order, forward)
is the reference's own code.  Seeded weights (`seeded.genh_state_dict`), a seeded 256 x 256 input, CPU fp32, eval mode;
stored: a strided sample + float64 moments of the output and of the tensor entering the decoder, and the state-dict key /
shape inventory.
"""
import json
import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402
from make_golden import GOLD, summarise  # noqa: E402
from megaportrait_hack_b200 import seeded  # noqa: E402


def main():
    ref = ref_shim.load_reference_model()
    RB = ref.ResBlock2D
    orig_init = RB.__init__

    def init_fixed(self, in_channels, out_channels=None, downsample=False):     # ResBlock2D(c) := ResBlock2D(c, c)
        orig_init(self, in_channels, in_channels if out_channels is None else out_channels, downsample)

    RB.__init__ = init_fixed
    try:
        G = ref.Genh().eval()
    finally:
        RB.__init__ = orig_init
    sd = seeded.genh_state_dict(0)
    missing = G.load_state_dict(sd, strict=True)
    print(missing)
    x = torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(5)) * 2 - 1
    mid = {}
    h = G.res_blocks.register_forward_hook(lambda m, i, o: mid.update(mid=o))
    with torch.no_grad():
        y = G(x)
    h.remove()
    blob = {}
    for k, v in (("out", y), ("mid", mid["mid"])):
        blob[k + ".sample"], blob[k + ".moments"] = summarise(v)
    np.savez_compressed(os.path.join(GOLD, "genh_synthetic.npz"), **blob)
    with open(os.path.join(GOLD, "genh_state_dict_keys.json"), "w") as f:
        json.dump({k: list(v.shape) for k, v in G.state_dict().items()}, f, indent=0)
    print("genh golden:", tuple(y.shape), float(y.abs().max()), tuple(mid["mid"].shape))


if __name__ == "__main__":
    main()
