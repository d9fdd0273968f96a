"""Row f-3: `Genh` / `GHR` (reference model.py:1346-1450, with the documented `ResBlock2D(64) ->
ResBlock2D(64, 64)` repair).  CPU: the oracle restatement against golden vectors minted from the REAL reference class
(oracle/make_golden_genh.py), state-dict inventory, host wiring on emulated kernels.  GPU (`-m gpu`): parity of the
libmpb200 path against the oracle."""
import json
import os

import numpy as np
import pytest
import torch

from conftest import GOLDEN, synthetic_pair

import gbase_oracle as O

TOL = 2e-5


def _genh_input():
    return torch.rand(1, 3, 256, 256, generator=torch.Generator().manual_seed(5)) * 2 - 1


@pytest.fixture(scope="module")
def genh_sd():
    from megaportrait_hack_b200 import seeded
    return seeded.genh_state_dict(0)


def test_genh_oracle_matches_reference_golden(genh_sd):
    gold = np.load(os.path.join(GOLDEN, "genh_synthetic.npz"))
    with torch.no_grad():
        y = O.genh(_genh_input(), genh_sd, "")
    sample, mom = gold["out.sample"], gold["out.moments"]
    flat = y.reshape(-1)
    assert flat.numel() == int(mom[4])
    err = np.abs(flat[:: int(mom[3])].numpy() - sample).max() / max(mom[2], 1e-6)
    assert err <= TOL, err
    assert abs(flat.double().pow(2).mean().sqrt().item() - mom[1]) <= TOL * mom[2]
    assert y.abs().max().item() < 1.0                      # tanh


def test_genh_state_dict_inventory_and_signatures(genh_sd):
    from megaportrait_hack_b200 import model
    want = json.load(open(os.path.join(GOLDEN, "genh_state_dict_keys.json")))
    G = model.Genh()
    got = {k: list(v.shape) for k, v in G.state_dict().items()}
    assert got == want                                      # the REAL reference class's keys and shapes
    assert set(genh_sd) == set(want)
    ghr = model.GHR()
    keys = list(ghr.state_dict())
    assert sum(k.startswith("Gbase.") for k in keys) == 971 and any(k.startswith("Genh.encoder.0.") for k in keys)
    # README.md:214 of the reference: GHR.Gbase.load_state_dict(Gbase.state_dict())
    ghr.Gbase.load_state_dict(model.Gbase().state_dict())
    import model as shim
    assert shim.Genh is model.Genh and shim.GHR is model.GHR


def test_genh_host_wiring_on_emulated_kernels(genh_sd):
    """Packing / BatchNorm folding / pooling and upsampling order of Genh._forward_cl on CPU with the kernels emulated."""
    import fake_ops
    from megaportrait_hack_b200 import model
    G = model.Genh().eval()
    G.load_state_dict(genh_sd)
    x = torch.rand(1, 3, 64, 64, generator=torch.Generator().manual_seed(6)) * 2 - 1
    with torch.no_grad():
        ref = O.genh(x, genh_sd, "")
        with fake_ops.installed():
            got = G(x)
    assert got.shape == ref.shape and (got - ref).abs().max().item() < 1e-4
    with pytest.raises(AssertionError):
        with fake_ops.installed():
            G(torch.zeros(1, 3, 60, 64))


@pytest.mark.gpu
def test_genh_parity_on_b200(genh_sd):
    from megaportrait_hack_b200 import lib, model, ops
    lib.build()
    G = model.Genh().eval()
    G.load_state_dict(genh_sd)
    G = G.cuda()
    x = torch.rand(2, 3, 512, 512, generator=torch.Generator().manual_seed(7)) * 2 - 1
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        ref = O.genh(x, genh_sd, "")
        l0 = ops.LAUNCHES
        got = G(x.cuda()).cpu()
    assert ops.LAUNCHES - l0 > 30, "Genh must run on libmpb200 kernels"
    err = (got - ref).abs().max().item()
    print("Genh 512x512 max-abs vs oracle:", err)
    assert err <= 1e-3
    # golden input (256 x 256) as well: the same tensor the real reference class produced
    gold = np.load(os.path.join(GOLDEN, "genh_synthetic.npz"))
    with torch.no_grad():
        y = G(_genh_input().cuda()).cpu().reshape(-1)
    assert np.abs(y[:: int(gold["out.moments"][3])].numpy() - gold["out.sample"]).max() <= 1e-3
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.Genh().eval()(torch.zeros(1, 3, 64, 64))


@pytest.mark.gpu
def test_ghr_end_to_end_on_b200(genh_sd, seeded_sd):
    """GHR.forward(xs, xd) = Genh(Gbase(xs, xd)[0]) (model.py:1446-1450) against the oracle chain."""
    import __graft_entry__ as entry
    from megaportrait_hack_b200 import model
    G, sd = entry.load_seeded_gbase("cuda")
    ghr = model.GHR().eval()
    ghr.Gbase = G
    ghr.Genh.load_state_dict(genh_sd)
    ghr = ghr.cuda()
    xs, xd = synthetic_pair(1)
    torch.set_num_threads(os.cpu_count() or 1)
    with torch.no_grad():
        ref = O.ghr_forward(xs, xd, sd, genh_sd)
        got = ghr(xs.cuda(), xd.cuda()).cpu()
    assert got.shape == (1, 3, 512, 512)
    err = (got - ref).abs().max().item()
    print("GHR max-abs vs oracle:", err)
    assert err <= 1e-3
