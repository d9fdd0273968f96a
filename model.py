"""Drop-in `model` module: `import model; model.Gbase()` resolves to the B200-native implementation
(megaportrait-hack_b200/model.py) under the reference's own names (reference model.py:54-1180).

Names of the reference's `model.py` that are outside the hot-path scope (losses, discriminator, student / high-res stages,
data utilities -- SURVEY.md section 2, rows 16-19; imported by train.py:16 and used as `model.Discriminator` at
train.py:418) are NOT re-implemented: they are delegated to the reference's own `model.py` when that file can be found
and imported (its optional dependencies -- lpips, facenet_pytorch, mediapipe, rembg, ... -- must be installed):

    MEGAPORTRAIT_REFERENCE=/path/to/MegaPortrait-hack   (default search: /root/reference, ./baseline/_ref)

The reference module is loaded under the private name `_mp_reference_model` (never as `model`), so the hot-path classes
above keep pointing at the B200 implementation; delegated classes that take a `Gbase` (PairwiseTransferLoss.forward,
model.py:2190-2219) work with the B200 `Gbase` because it keeps the reference's attribute names.  When the reference
cannot be imported the names raise an ImportError that says why.
"""
import importlib.util as _ilu
import os as _os
import sys as _sys

from megaportrait_hack_b200.model import *  # noqa: F401,F403
from megaportrait_hack_b200.model import (  # noqa: F401
    COMPRESS_DIM, FEATURE_SIZE, FEATURE_SIZE_AVG_POOL, AdaptiveGroupNorm, AntiAliasInterpolation2d, Conv2d_WS,
    Conv3D_WS, CustomResNet50, Eapp, Emtn, FlowField, G2d, G3d, Gbase, Genh, GHR, ImagePyramide, ResBlock2D, ResBlock3D,
    ResBlock3D_Adaptive, ResBlock_Custom, SixDRepNet_Detector, WarpGeneratorC2D, WarpGeneratorS2C,
    apply_warping_field, compute_rotation_matrix, compute_rt_warp, device, invalidate_plans, load_reference_state_dict)

_OUT_OF_SCOPE = ("PerceptualLoss", "IdentitySimilarityLoss", "PairwiseTransferLoss", "Discriminator",
                 "Student", "crop_and_warp_face", "get_foreground_mask", "remove_background_and_convert_to_rgb",
                 "GazeBlinkLoss", "MPGazeLoss", "PatchGanEncoder", "GazeLoss", "Encoder", "Decoder", "UNet",
                 "cosine_loss", "contrastive_loss", "ResBlock", "PatchDiscriminator", "MultiscaleDiscriminator")
_ref_state = {"module": None, "error": None}


def _reference_dirs():
    here = _os.path.dirname(_os.path.abspath(__file__))
    env = _os.environ.get("MEGAPORTRAIT_REFERENCE")
    return ([env] if env else []) + ["/root/reference", _os.path.join(here, "baseline", "_ref")]


def _load_reference():
    """Import the reference's model.py as `_mp_reference_model` (once)."""
    if _ref_state["module"] is not None or _ref_state["error"] is not None:
        return _ref_state["module"]
    here = _os.path.abspath(__file__)
    for d in _reference_dirs():
        path = _os.path.join(d, "model.py")
        if not _os.path.isfile(path) or _os.path.abspath(path) == here:
            continue
        spec = _ilu.spec_from_file_location("_mp_reference_model", path)
        mod = _ilu.module_from_spec(spec)
        added = d not in _sys.path
        if added:
            _sys.path.append(d)      # the reference imports its siblings (resnet.py, mysixdrepnet.py, ...); appended AFTER this shim
        try:
            _sys.modules["_mp_reference_model"] = mod
            spec.loader.exec_module(mod)
            _ref_state["module"] = mod
            return mod
        except Exception as e:       # a missing optional dependency of the reference, usually
            _sys.modules.pop("_mp_reference_model", None)
            _ref_state["error"] = f"{path}: {type(e).__name__}: {e}"
            if added:
                _sys.path.remove(d)
    if _ref_state["error"] is None:
        _ref_state["error"] = "no reference model.py found in " + ", ".join(_reference_dirs())
    return None


def __getattr__(name):
    if name.startswith("__"):
        raise AttributeError(name)
    ref = _load_reference() if (name in _OUT_OF_SCOPE or not name.startswith("_")) else None
    if ref is not None and hasattr(ref, name):
        return getattr(ref, name)
    if name in _OUT_OF_SCOPE:
        raise ImportError(f"model.{name} is outside the B200 hot-path scope (SURVEY.md section 2/8f) and is delegated to "
                          f"the reference's own model.py, which could not be imported here ({_ref_state['error']}); set "
                          "MEGAPORTRAIT_REFERENCE to a checkout whose optional dependencies are installed")
    raise AttributeError(f"module 'model' has no attribute {name!r}")
