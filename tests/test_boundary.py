"""Drop-in boundary: names, signatures, state_dict inventory, C-ABI exports.  CPU only."""
import ctypes
import inspect
import json
import os
import re

import pytest
import torch

from conftest import GOLDEN, ROOT


@pytest.fixture(scope="module")
def model():
    import model as m   # the root-level drop-in shim
    return m


def test_state_dict_inventory_matches_reference(model):
    ref = json.load(open(os.path.join(GOLDEN, "state_dict_keys.json")))
    sd = model.Gbase().state_dict()
    assert len(ref) == 971
    assert set(sd) == set(ref)
    assert all(list(sd[k].shape) == ref[k] for k in ref)
    assert list(sd)[:3] == list(ref)[:3]   # same registration order


def test_reference_state_dict_loads_strict(model, seeded_sd):
    from megaportrait_hack_b200 import seeded
    G = model.Gbase()
    sd = {k: v for k, v in seeded_sd.items() if not k.startswith(seeded.ROTNET_PREFIX)}
    sd.update({k: v for k, v in G.state_dict().items() if k.startswith("image_pyramid.")})
    G.load_state_dict(sd, strict=True)
    G.motionEncoder.rotation_net.model.load_state_dict(
        {k[len(seeded.ROTNET_PREFIX):]: v for k, v in seeded_sd.items() if k.startswith(seeded.ROTNET_PREFIX)},
        strict=True)


def test_signatures_match_reference(model):
    sigs = json.load(open(os.path.join(GOLDEN, "signatures.json")))
    norm = lambda s: re.sub(r"\s+", "", s)
    for name, d in sigs.items():
        obj = getattr(model, name)
        if "call" in d:
            assert norm(str(inspect.signature(obj))) == norm(d["call"]), name
        else:
            assert norm(str(inspect.signature(obj.forward))) == norm(d["forward"]), name
            ref_init = norm(d["init"])
            got_init = norm(str(inspect.signature(obj.__init__)))
            if name in ("Conv2d_WS", "Conv3D_WS", "CustomResNet50"):
                continue   # inherit nn.Conv*/torchvision ctor signatures
            assert got_init == ref_init, f"{name}: {got_init} != {ref_init}"


def test_gbase_attributes_and_out_of_scope_names(model):
    G = model.Gbase()
    for a in ("appearanceEncoder", "motionEncoder", "warp_generator_s2c", "warp_generator_c2d", "G3d", "G2d",
              "image_pyramid"):
        assert hasattr(G, a)
    assert G.training   # default mode after construction is train(), as in the reference
    with pytest.raises(ImportError):
        model.PerceptualLoss
    k05 = G.image_pyramid.downs["0-5"].weight
    assert k05.shape == (3, 1, 5, 5) and abs(k05[0].sum().item() - 1.0) < 1e-6


def test_cpu_inputs_raise_not_fallback(model):
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.apply_warping_field(torch.zeros(1, 2, 4, 4, 4), torch.zeros(1, 3, 4, 4, 4))
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        model.G3d(96).eval()(torch.zeros(1, 96, 16, 64, 64))


def test_c_abi_library_exports_every_declared_symbol():
    from megaportrait_hack_b200 import lib
    lib.build()
    header = open(os.path.join(ROOT, "include", "mpb200.h")).read()
    declared = set(re.findall(r"\b(mp_[a-z0-9_]+)\s*\(", header)) - {"mp_conv_desc"}
    assert len(declared) >= 20
    so = ctypes.CDLL(lib.LIB_PATH)
    for name in declared:
        assert hasattr(so, name), f"libmpb200.so does not export {name}"
    assert set(lib.EXPORTS) == declared
    assert lib.load().mp_abi_version() == lib.ABI_VERSION == 4
    # 12 pointers, 22 ints, 2 pointers, 2 ints + 3 floats (+ pad to 8)
    assert ctypes.sizeof(lib.ConvDesc) == 12 * 8 + 22 * 4 + 2 * 8 + 6 * 4
