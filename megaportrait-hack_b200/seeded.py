"""Deterministic synthetic weights for Gbase (there is no checkpoint in the reference tree, and no network).

The reference's own random init (`torch.manual_seed(0); model.Gbase()`) depends on the construction order of
~400 sub-modules and cannot be reproduced outside the reference, so every tensor is generated from its
*state_dict key* instead: the same recipe fills the real reference (golden-vector generation,
`oracle/make_golden.py`), the CPU oracle (`oracle/gbase_oracle.py`) and the B200 modules, on any box.

Distributions follow PyTorch's default inits (`nn.Conv*`/`nn.Linear`: U(-1/sqrt(fan_in), 1/sqrt(fan_in)),
`torch.randn` for `adaptive_matrix_*`, model.py:934-935) except that normalisation scales/shifts and the
BatchNorm running statistics are perturbed away from (1, 0, 0, 1) so that BN folding, the double affine of
`AdaptiveGroupNorm` (model.py:314-316) and GroupNorm affine paths are actually exercised by parity tests.
"""
from __future__ import annotations

import json
import math
import os
import zlib
from typing import Dict, Iterable, Tuple

import torch

# kinds
CONV_W, CONV_B, NORM_W, NORM_B, BN_MEAN, BN_VAR, COUNT, RANDN, KEEP = (
    "conv_w", "conv_b", "norm_w", "norm_b", "bn_mean", "bn_var", "count", "randn", "keep")

ROTNET_PREFIX = "motionEncoder.rotation_net.model."  # 6DRepNet lives outside state_dict() (mysixdrepnet.py:771)
WEIGHT_GAIN = 1.0


def _gen(key: str, seed: int) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed(((zlib.crc32(key.encode()) + 0x9E3779B1 * (seed + 1)) & 0x7FFFFFFFFFFF))
    return g


def seeded_tensor(key: str, shape: Tuple[int, ...], kind: str, seed: int = 0) -> torch.Tensor:
    g = _gen(key, seed)
    shape = tuple(shape)
    if kind == CONV_W:
        fan_in = 1
        for s in shape[1:]:
            fan_in *= s
        b = WEIGHT_GAIN / math.sqrt(fan_in)
        return (torch.rand(shape, generator=g) * 2 - 1) * b
    if kind == CONV_B:
        return (torch.rand(shape, generator=g) * 2 - 1) * 0.05
    if kind == NORM_W:
        return 0.8 + 0.4 * torch.rand(shape, generator=g)
    if kind == NORM_B:
        return (torch.rand(shape, generator=g) * 2 - 1) * 0.1
    if kind == BN_MEAN:
        return (torch.rand(shape, generator=g) * 2 - 1) * 0.1
    if kind == BN_VAR:
        return 0.8 + 0.4 * torch.rand(shape, generator=g)
    if kind == COUNT:
        return torch.zeros(shape, dtype=torch.long)
    if kind == RANDN:
        return torch.randn(shape, generator=g)
    raise ValueError(f"no recipe for kind {kind!r} ({key})")


def classify_module_tensors(model: torch.nn.Module) -> Dict[str, Tuple[Tuple[int, ...], str]]:
    """key -> (shape, kind) for every entry of `model.state_dict()`, derived from the owning module's type."""
    nn = torch.nn
    out: Dict[str, Tuple[Tuple[int, ...], str]] = {}
    for mname, mod in model.named_modules():
        prefix = mname + "." if mname else ""
        direct = list(mod.named_parameters(recurse=False)) + list(mod.named_buffers(recurse=False))
        for pname, t in direct:
            key = prefix + pname
            shape = tuple(t.shape)
            if key.startswith("image_pyramid."):
                kind = KEEP  # Gaussian kernels computed by the ctor (model.py:652-679)
            elif isinstance(mod, (nn.Conv1d, nn.Conv2d, nn.Conv3d, nn.Linear)):
                kind = CONV_W if pname == "weight" else CONV_B
            elif isinstance(mod, (nn.BatchNorm1d, nn.BatchNorm2d, nn.BatchNorm3d)):
                kind = {"weight": NORM_W, "bias": NORM_B, "running_mean": BN_MEAN, "running_var": BN_VAR,
                        "num_batches_tracked": COUNT}[pname]
            elif isinstance(mod, nn.GroupNorm) or type(mod).__name__ == "AdaptiveGroupNorm":
                kind = NORM_W if pname == "weight" else NORM_B
            elif pname.startswith("adaptive_matrix_"):
                kind = RANDN
            else:
                raise ValueError(f"unclassified tensor {key} in {type(mod).__name__}")
            out[key] = (shape, kind)
    return out


def manifest_path() -> str:
    here = os.path.dirname(os.path.abspath(__file__))
    return os.path.join(here, "gbase_manifest.json")


def load_manifest() -> Dict[str, Tuple[Tuple[int, ...], str]]:
    """The reference's 971 state_dict entries (+58 un-registered 6DRepNet tensors): key -> (shape, kind)."""
    with open(manifest_path()) as f:
        raw = json.load(f)
    return {k: (tuple(v[0]), v[1]) for k, v in raw.items()}


def seeded_state_dict(manifest: Dict[str, Tuple[Tuple[int, ...], str]] | None = None, seed: int = 0,
                      keys: Iterable[str] | None = None) -> Dict[str, torch.Tensor]:
    manifest = manifest or load_manifest()
    sd = {}
    for k in (keys if keys is not None else manifest.keys()):
        shape, kind = manifest[k]
        if kind == KEEP:
            continue
        sd[k] = seeded_tensor(k, shape, kind, seed)
    return sd


def apply_seeded(model: torch.nn.Module, seed: int = 0, rotnet: torch.nn.Module | None = None) -> None:
    """Fill `model` (reference Gbase or ours) in place from the per-key recipe."""
    cls = classify_module_tensors(model)
    sd = model.state_dict()
    with torch.no_grad():
        for k, (shape, kind) in cls.items():
            if kind == KEEP:
                continue
            sd[k].copy_(seeded_tensor(k, shape, kind, seed))
        if rotnet is not None:
            rcls = classify_module_tensors(rotnet)
            rsd = rotnet.state_dict()
            for k, (shape, kind) in rcls.items():
                rsd[k].copy_(seeded_tensor(ROTNET_PREFIX + k, shape, kind, seed))


GENH_PREFIX = "Genh."


def genh_state_dict(seed: int = 0) -> Dict[str, torch.Tensor]:
    """Seeded weights of `Genh`, keyed like `Genh().state_dict()`; the generator of every tensor is
    derived from "Genh." + key, so the reference's Genh, the oracle and the B200 module get the same values."""
    from . import model
    cls = classify_module_tensors(model.Genh())
    return {k: seeded_tensor(GENH_PREFIX + k, shape, kind, seed) for k, (shape, kind) in cls.items()}
