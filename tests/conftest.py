import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
for p in (ROOT, os.path.join(ROOT, "oracle")):
    if p not in sys.path:
        sys.path.insert(0, p)

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with `-m gpu`)")


def pytest_collection_modifyitems(config, items):
    try:
        import torch
        has_gpu = torch.cuda.is_available()
    except Exception:
        has_gpu = False
    if has_gpu:
        return
    skip = pytest.mark.skip(reason="no CUDA device")
    for it in items:
        if "gpu" in it.keywords:
            it.add_marker(skip)


@pytest.fixture(scope="session")
def seeded_sd():
    from megaportrait_hack_b200 import seeded
    return seeded.seeded_state_dict(seed=0)


def load_frames():
    """The two real 512x512 frames of BASELINE config 1 (fixtures minted by oracle/make_golden.py)."""
    import cv2
    import torch
    out = []
    for tag in ("src", "drv"):
        bgr = cv2.imread(os.path.join(GOLDEN, f"frame_{tag}.png"))
        rgb = cv2.cvtColor(bgr, cv2.COLOR_BGR2RGB)
        out.append(torch.from_numpy(rgb).permute(2, 0, 1)[None].float() / 255.0)
    return out


def synthetic_pair(n_drv=1):
    import torch
    g = torch.Generator().manual_seed(1)
    xs = torch.rand(1, 3, 512, 512, generator=g)
    xd = torch.rand(n_drv, 3, 512, 512, generator=g)
    return xs, xd
