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


def test_out_of_scope_names_are_delegated_to_the_reference(tmp_path):
    """SURVEY.md 8b / VERDICT round 1 missing #4: `from model import PerceptualLoss, ...` (train.py:16) and
    `model.Discriminator()` (train.py:418) resolve to the reference's own model.py when it is importable; the hot-path
    classes keep resolving to the B200 implementation; without an importable reference the names raise ImportError."""
    import subprocess
    import sys
    ref = tmp_path / "ref"
    ref.mkdir()
    (ref / "model.py").write_text(
        "import torch.nn as nn\n"
        "class Gbase(nn.Module):\n    pass\n"                       # must NOT shadow the B200 Gbase
        "class Discriminator(nn.Module):\n    def forward(self, x):\n        return x * 2\n"
        "class PerceptualLoss(nn.Module):\n    pass\n"
        "def crop_and_warp_face(x):\n    return 'ref:' + str(x)\n")
    code = ("import model, torch\n"
            "from model import PerceptualLoss, crop_and_warp_face, apply_warping_field\n"
            "assert model.Discriminator()(torch.ones(1)).item() == 2.0\n"
            "assert crop_and_warp_face(3) == 'ref:3'\n"
            "assert model.Gbase.__module__.endswith('megaportrait_hack_b200.model'), model.Gbase.__module__\n"
            "assert PerceptualLoss.__module__ == '_mp_reference_model'\n"
            "try:\n    model.IdentitySimilarityLoss\n    raise SystemExit('expected ImportError')\n"
            "except ImportError as e:\n    assert 'delegated' in str(e)\n"
            "print('DELEGATION_OK')\n")
    env = dict(os.environ, MEGAPORTRAIT_REFERENCE=str(ref))
    r = subprocess.run([sys.executable, "-c", code], cwd=ROOT, env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "DELEGATION_OK" in r.stdout, r.stdout[-2000:] + r.stderr[-2000:]
    # no importable reference: a clear ImportError (in this container /root/reference exists but lacks its optional deps)
    code2 = ("import model\n"
             "try:\n    model.Discriminator\n    print('HAS_REF')\n"
             "except ImportError as e:\n    assert 'could not be imported' in str(e); print('IMPORT_ERROR_OK')\n")
    env2 = dict(os.environ, MEGAPORTRAIT_REFERENCE=str(tmp_path / "missing"))
    r2 = subprocess.run([sys.executable, "-c", code2], cwd=ROOT, env=env2, capture_output=True, text=True, timeout=300)
    assert r2.returncode == 0 and ("IMPORT_ERROR_OK" in r2.stdout or "HAS_REF" in r2.stdout), r2.stdout + r2.stderr[-2000:]


def test_reference_checkpoint_loader_tolerates_the_cuda_build_quirk():
    """a checkpoint written by a CUDA build of the reference lacks the 4 `adaptive_matrix_*` keys
    (model.py:934-935); `load_reference_state_dict` accepts exactly that, warns, and rejects anything else."""
    import warnings
    import model
    G = model.Gbase()
    sd = {k: v.clone() for k, v in G.state_dict().items()}
    cuda_style = {k: v for k, v in sd.items() if "adaptive_matrix_" not in k}
    assert len(sd) - len(cuda_style) == 4
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        model.load_reference_state_dict(G, cuda_style)
    assert any("adaptive_matrix" in str(x.message) for x in w)
    mats = {k: torch.full_like(sd[k], 0.25) for k in sd if "adaptive_matrix_" in k}
    with warnings.catch_warnings(record=True) as w:
        warnings.simplefilter("always")
        model.load_reference_state_dict(G, cuda_style, adaptive_matrices=mats)
    assert not w and float(G.warp_generator_s2c.adaptive_matrix_gamma[0, 0]) == 0.25
    broken = dict(cuda_style)
    broken.pop("G3d.final_conv.weight")
    with pytest.raises(RuntimeError, match="missing"):
        model.load_reference_state_dict(G, broken)


def test_sharded_step_with_more_ranks_than_frames_returns_empty_outputs():
    from megaportrait_hack_b200 import engine

    class _G:
        def encode_source(self, xs):
            return {"vc2d": None, "es": None}

        def drive(self, src, xd):
            raise AssertionError("drive must not run on an empty shard")

    out, pyr = engine.ShardedGbase(_G()).step(torch.zeros(1, 3, 512, 512), torch.zeros(0, 3, 512, 512))
    assert out.shape == (0, 3, 512, 512) and pyr["prediction_0.25"].shape == (0, 3, 128, 128)
