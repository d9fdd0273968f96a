"""TEST INFRASTRUCTURE ONLY -- import the *real* reference (`/root/reference/model.py`) in this container.

The reference cannot be imported unmodified: `model.py:7-34` imports ten optional packages that are not
installed, `mysixdrepnet.py:903` needs a numpy<2 private symbol, `resnet.py:289-300` and
`mysixdrepnet.py:792` download weights, and every ctor does `.to(device)` with the import-time global
(`model.py:51`).  None of that is on the Gbase hot path, so this shim stubs it.

Used only by `oracle/make_golden.py` (golden-vector generation) and by CPU tests that are skipped when
`/root/reference` is absent (it does not exist on the GPU box).  Nothing in the product imports this file.
"""
from __future__ import annotations

import os
import sys
import types

REFERENCE_DIR = os.environ.get("MP_REFERENCE_DIR", "/root/reference")

_STUBS = [
    "colored_traceback", "colored_traceback.auto", "torchsummary", "memory_profiler", "facenet_pytorch",
    "skimage", "skimage.transform", "face_recognition", "lpips", "mediapipe", "rembg",
    "matplotlib", "matplotlib.pyplot", "matplotlib.patches", "mpl_toolkits", "mpl_toolkits.mplot3d",
]


def reference_available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_DIR, "model.py"))


class _Dummy:
    def __init__(self, *a, **k):
        pass

    def __call__(self, *a, **k):
        return self

    def __getattr__(self, name):
        if name.startswith("__"):
            raise AttributeError(name)
        return _Dummy()


def _install_stubs() -> None:
    for name in _STUBS:
        if name in sys.modules:
            continue
        try:
            __import__(name)
            continue
        except Exception:
            pass
        m = types.ModuleType(name)
        m.__path__ = []  # behave as a package so sub-imports resolve
        def _ga(attr, _n=name):
            if attr.startswith("__"):
                raise AttributeError(attr)
            return _Dummy()
        m.__getattr__ = _ga  # type: ignore[attr-defined]
        sys.modules[name] = m
    sys.modules["memory_profiler"].profile = lambda f=None, **k: f if f is not None else (lambda g: g)
    sys.modules["torchsummary"].summary = lambda *a, **k: None
    import numpy.lib as _nl
    if "numpy.lib.function_base" not in sys.modules:
        fb = types.ModuleType("numpy.lib.function_base")
        fb._quantile_unchecked = None
        sys.modules["numpy.lib.function_base"] = fb
        _nl.function_base = fb


def load_reference_model():
    """Returns the reference's `model` module, patched so `model.Gbase()` builds offline on CPU."""
    if not reference_available():
        raise FileNotFoundError(f"reference not found at {REFERENCE_DIR}")
    if "model" in sys.modules and getattr(sys.modules["model"], "__mp_ref_shim__", False):
        return sys.modules["model"]
    _install_stubs()
    saved = sys.modules.pop("model", None)
    sys.path.insert(0, REFERENCE_DIR)
    try:
        import torch
        import importlib
        ref = importlib.import_module("model")
    finally:
        sys.path.remove(REFERENCE_DIR)
    ref.__mp_ref_shim__ = True
    ref.device = torch.device("cpu")
    import mysixdrepnet

    _orig_resnet18 = ref.resnet18

    def _resnet18_offline(pretrained=False, **kw):
        return _orig_resnet18(pretrained=False, **kw)

    ref.resnet18 = _resnet18_offline

    class _OfflineSixDRepNet(mysixdrepnet.SixDRepNet_Detector):
        def __init__(self, gpu_id=-1, dict_path=""):
            self.gpu = -1
            self.model = mysixdrepnet.MySixDRepNet("RepVGG-B1g2", "", deploy=True, pretrained=False).eval()

    ref.SixDRepNet_Detector = _OfflineSixDRepNet
    # keep the reference importable under its own name and ours
    sys.modules["mp_reference_model"] = ref
    if saved is not None:
        sys.modules["model"] = saved
    return ref
