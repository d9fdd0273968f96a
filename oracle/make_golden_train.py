"""TEST INFRASTRUCTURE ONLY -- golden GRADIENTS from the REAL reference in train mode.

Run in the authoring container (the reference is not on the GPU box):

    python oracle/make_golden_train.py

Imports `/root/reference/model.py` through `oracle/ref_shim.py`, builds `Gbase()`, fills it with the per-key seeded weights
(`megaportrait_hack_b200/seeded.py`, seed 0), puts it in `.train()` mode exactly as `train.py:133` does (BatchNorm: batch
statistics; the un-registered 6DRepNet stays in eval mode, as in the reference), runs `Gbase(xs, xd)` on the synthetic pair of
SURVEY.md 8d on CPU fp32 and back-propagates `rgb.mean() + pyr_0.5.mean() + pyr_0.25.mean()`.  For every registered parameter it
stores the gradient's float64 L2 norm and a strided sample (<= 256 values) in `tests/golden/gbase_train.npz`, plus a sample of the
train-mode image.  `tests/test_oracle_golden.py::test_oracle_train_mode_matches_reference_golden` holds
`oracle/gbase_oracle.py` (`BN_TRAINING`, `gbase_forward_train`) to these: that is what pins the checker of the CUDA training path.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path.insert(0, ROOT)
sys.path.insert(0, HERE)

import ref_shim  # noqa: E402
from megaportrait_hack_b200 import seeded  # noqa: E402

GOLD = os.path.join(ROOT, "tests", "golden")
MAX_SAMPLES = 256


def sample(t: torch.Tensor):
    flat = t.detach().reshape(-1).to(torch.float32)
    stride = max(1, (flat.numel() + MAX_SAMPLES - 1) // MAX_SAMPLES)
    return flat[::stride].numpy().copy()


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    ref = ref_shim.load_reference_model()
    G = ref.Gbase()
    seeded.apply_seeded(G, 0, rotnet=G.motionEncoder.rotation_net.model)
    G.train()
    g = torch.Generator().manual_seed(1)
    xs = torch.rand(1, 3, 512, 512, generator=g)
    xd = torch.rand(1, 3, 512, 512, generator=g)
    rgb, pyr = G(xs, xd)
    loss = rgb.mean() + pyr["prediction_0.5"].mean() + pyr["prediction_0.25"].mean()
    loss.backward()
    blob = {"rgb.sample": sample(rgb), "loss": np.array([float(loss)])}
    names = []
    for name, p in G.named_parameters():
        if p.grad is None:
            continue
        names.append(name)
        blob["grad." + name + ".sample"] = sample(p.grad)
        blob["grad." + name + ".norm"] = np.array([float(p.grad.double().norm())])
    blob["names"] = np.array(names)
    np.savez_compressed(os.path.join(GOLD, "gbase_train.npz"), **blob)
    print(f"loss {float(loss):.8f}; {len(names)} parameter gradients stored")


if __name__ == "__main__":
    main()
