"""Drop-in replacements for the hot-path classes of the reference's `model.py` (johndpope/MegaPortrait-hack).

Same class names, constructor arguments, `forward()` signatures, attribute names and `state_dict()` keys as the
reference (model.py:54-1180), so `train.py` / `inference.py` / `PairwiseTransferLoss` call them unchanged; the
arithmetic runs in the sm_100a kernels of libmpb200 (include/mpb200.h).  There is NO CPU path and NO ATen fallback
for the hot-path operators: CPU tensors raise.  Under `torch.no_grad()` in eval mode the modules run the fused inference
kernels; in `.train()` mode or with autograd recording they run the differentiable libmpb200 path (`_forward_autograd`,
SURVEY.md row f-2: every convolution in all three directions, GroupNorm, train-mode BatchNorm and both warps on the GPU kernels).

Two calling levels:
  * module level -- every class takes / returns the reference's NCHW / NCDHW fp32 tensors;
  * fused level  -- `Gbase.forward`, `Gbase.encode_source`, `Gbase.drive` chain the stages on channels-last
    split-bf16 activations without leaving the internal layout.
"""
from __future__ import annotations

import math
from typing import Dict, Optional, Tuple

import os

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .emtn import COMPRESS_DIM, FEATURE_SIZE, FEATURE_SIZE_AVG_POOL, CustomResNet50, Emtn, SixDRepNet_Detector  # noqa: F401
from .ops import ACT_NONE, ACT_RELU, ACT_RELU_TANH, ACT_SIGMOID, ACT_TANH, Act

# G2d's identity res-blocks on the fp16 + FP8 cross-term convolution (MPB200_G2D_PREC=split selects three-pass split-bf16)
_Q8_ENABLED = os.environ.get("MPB200_G2D_PREC", "q8") != "split"
# `Gbase.forward` replays a CUDA graph per (batch size, device) once a shape has been seen twice (MPB200_FORWARD_GRAPHS=0: always
# launch eagerly); the attribute `Gbase.forward_graphs` switches it per instance
_FWD_GRAPHS = os.environ.get("MPB200_FORWARD_GRAPHS", "1") != "0"
# ... and its first two up-blocks (MPB200_G2D_UP_PREC=split keeps them on three-pass split-bf16)
_Q8_UP = os.environ.get("MPB200_G2D_UP_PREC", "q8") != "split"
# source half at small batch: the three branches that only read the frame run on three streams (0 = one stream; A/B runs)
_SOURCE_FORK = os.environ.get("MPB200_SOURCE_FORK", "1") != "0"
_SIDE_STREAMS: Dict[object, Tuple["torch.cuda.Stream", "torch.cuda.Stream"]] = {}


def _side_streams(device):
    key = torch.device(device)
    if key not in _SIDE_STREAMS:
        _SIDE_STREAMS[key] = (torch.cuda.Stream(key), torch.cuda.Stream(key))
    return _SIDE_STREAMS[key]

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")   # model.py:51 (kept for callers that read it)


# ----------------------------------------------------------------------------------------------------- guards
def _require_inference(mod: nn.Module, *tensors) -> None:
    for t in tensors:
        if torch.is_tensor(t) and not t.is_cuda:
            raise RuntimeError(f"{type(mod).__name__}: input is on {t.device}; the B200 path has no CPU fallback "
                               "(move the module and its inputs to a CUDA device)")
    if torch.is_grad_enabled() and any(torch.is_tensor(t) and t.requires_grad for t in tensors):
        raise NotImplementedError(f"{type(mod).__name__}: backward is not implemented (SURVEY.md 8f-2); "
                                  "call under torch.no_grad()")


def _wants_grad(mod: nn.Module, *tensors) -> bool:
    """Autograd is recording and an input or a parameter of `mod` asks for a gradient: modules that have a differentiable
    libmpb200 path (row f-2: ResBlock3D, G3d) take it; the others raise."""
    return torch.is_grad_enabled() and (any(torch.is_tensor(t) and t.requires_grad for t in tensors)
                                        or any(p.requires_grad for p in mod.parameters()))


def _public(y):
    """Outputs of the differentiable path are channels-last in memory (ops._cl_view); a module's public `forward()` hands its
    callers the reference's contiguous NCHW / NCDHW layout (`.view()` on it keeps working)."""
    if torch.is_tensor(y):
        return y.contiguous()
    if isinstance(y, dict):
        return {k: _public(v) for k, v in y.items()}
    return tuple(_public(v) for v in y)


def _sig(mod: nn.Module):
    """Change-detector for cached packed weights: per-tensor (data_ptr, version) of every parameter and buffer, the
    device and the train flag.  In-place writes through `param.data` / `.data.copy_()` do NOT bump `_version`: after
    such writes (EMA updates, hand-written checkpoint loaders) call `invalidate_plans(module)`."""
    ts = list(mod.parameters()) + list(mod.buffers())
    return (tuple((t.data_ptr(), t._version) for t in ts), str(ts[0].device) if ts else "", mod.training)


def invalidate_plans(root: nn.Module) -> None:
    """Drop every cached kernel-format weight plan under `root` (they are rebuilt on the next forward)."""
    for m in root.modules():
        for k in ("_mp_plan", "_mp_final", "_mp_plans", "_mp_cuda_plans", "_mp_cuda_plan", "_mp_cuda_rot_plan", "_mp_fwd_graphs",
                  "_mp_fwd_seen"):
            m.__dict__.pop(k, None)
        det = getattr(m, "rotation_net", None)
        if det is not None and hasattr(det, "model"):
            det.model.__dict__.pop("_mp_plan", None)


class _Packed:
    """Mixin: lazily (re)built kernel-format weights, keyed on parameter versions."""

    def _plan(self):
        sig = _sig(self)
        cache = self.__dict__.get("_mp_plan")
        if cache is None or cache[0] != sig:
            with torch.no_grad():
                cache = (sig, self._build_plan())
            self.__dict__["_mp_plan"] = cache
        return cache[1]

    def _load_from_state_dict(self, *a, **k):      # checkpoint loads copy through `.data`-like paths: drop the plan
        self.__dict__.pop("_mp_plan", None)
        return super()._load_from_state_dict(*a, **k)


def _capturing(a) -> bool:
    """True while the current CUDA stream is being captured into a graph (host syncs are illegal there)."""
    return a.device.type == "cuda" and torch.cuda.is_current_stream_capturing()


def _f32c(t: torch.Tensor) -> torch.Tensor:
    return t.detach().float().contiguous()


def _as_f32_cuda(x: torch.Tensor) -> torch.Tensor:
    return x.detach().to(torch.float32).contiguous()


# ----------------------------------------------------------------------------------------------------- layers
class Conv2d_WS(nn.Conv2d):
    """model.py:54-69: weight-standardised conv."""

    def _forward_autograd(self, x, stats_groups: int = 0):
        """Differentiable form (row f-2): the standardisation is a float64 torch expression on the weight (autograd carries its
        Jacobian), the convolution and its three gradients run on libmpb200 (ops.ConvFunction).  `stats_groups`: also return
        the GroupNorm statistics of the output from the convolution's epilogue -> (y, stats)."""
        return ops.conv_train(x, ops.standardize_weight(self.weight).float(), self.bias, stats_groups=stats_groups)

    def forward(self, x):
        if _wants_grad(self, x):
            _require_inference(self, x.detach())
            return _public(self._forward_autograd(x.float()))
        _require_inference(self, x)
        pw = ops.pack_conv(ops.standardize_weight(self.weight), self.bias, x.device)
        a = ops.from_nchw(_as_f32_cuda(x))
        out, _ = ops.conv(a, pw)
        return ops.to_nchw(out, 4)


class Conv3D_WS(nn.Conv3d):
    """model.py:71-86."""

    def _forward_autograd(self, x):
        return ops.ConvFunction.apply(x, ops.standardize_weight(self.weight).float(), self.bias)

    def forward(self, x):
        if _wants_grad(self, x):
            _require_inference(self, x.detach())
            return _public(self._forward_autograd(x.float()))
        _require_inference(self, x)
        pw = ops.pack_conv(ops.standardize_weight(self.weight), self.bias, x.device)
        a = ops.from_nchw(_as_f32_cuda(x))
        out, _ = ops.conv(a, pw)
        return ops.to_nchw(out, 5)


class ResBlock_Custom(nn.Module, _Packed):
    """model.py:88-130: conv_res(x) + conv(relu(gn(conv_ws(relu(gn(x)))))), GN(32) non-affine."""

    def __init__(self, dimension, in_channels, out_channels):
        super().__init__()
        self.dimension = dimension
        self.in_channels = in_channels
        self.out_channels = out_channels
        if dimension == 2:
            self.conv_res = nn.Conv2d(in_channels, out_channels, 3, padding=1)
            self.conv_ws = Conv2d_WS(in_channels=in_channels, out_channels=out_channels, kernel_size=3, padding=1)
            self.conv = nn.Conv2d(out_channels, out_channels, 3, padding=1)
        elif dimension == 3:
            self.conv_res = nn.Conv3d(in_channels, out_channels, 3, padding=1)
            self.conv_ws = Conv3D_WS(in_channels=in_channels, out_channels=out_channels, kernel_size=3, padding=1)
            self.conv = nn.Conv3d(out_channels, out_channels, 3, padding=1)

    def _build_plan(self):
        dev = self.conv.weight.device
        return {"res": ops.pack_conv(self.conv_res.weight, self.conv_res.bias, dev),
                "ws": ops.pack_conv(ops.standardize_weight(self.conv_ws.weight), self.conv_ws.bias, dev),
                "conv": ops.pack_conv(self.conv.weight, self.conv.bias, dev)}

    def _forward_cl(self, x: Act, x_stats: Optional[torch.Tensor]) -> Act:
        """x needs f32 + split; returns f32 (the sum out1 + out2)."""
        P = self._plan()
        out2, _ = ops.conv(x, P["res"], f32=True)
        h = ops.group_norm_act(x, 32, x_stats, act=ACT_RELU, split=True)
        h, st = ops.conv(h, P["ws"], f32=True, stats_groups=32)
        h = ops.group_norm_act(h, 32, st, act=ACT_RELU, split=True)
        out, _ = ops.conv(h, P["conv"], res=out2, f32=True)
        return out

    def _forward_autograd(self, x):
        """Differentiable form (row f-2): ops.ConvFunction / ops.GroupNormFunction (CUDA forward and backward)."""
        conv, gn = ops.ConvFunction.apply, ops.GroupNormFunction.apply
        out2 = conv(x, self.conv_res.weight, self.conv_res.bias)
        h = torch.relu(gn(x, 32, None, None, 1e-5))
        t, st = self.conv_ws._forward_autograd(h, stats_groups=32)       # GroupNorm statistics from the convolution's epilogue
        h = torch.relu(gn(t, 32, None, None, 1e-5, st))
        return conv(h, self.conv.weight, self.conv.bias) + out2

    def forward(self, x):
        if _wants_grad(self, x):
            _require_inference(self, x.detach())
            return _public(self._forward_autograd(x.float()))
        _require_inference(self, x)
        a = ops.from_nchw(_as_f32_cuda(x), f32=True, split=True)
        out = self._forward_cl(a, None)
        y = ops.to_nchw(out, x.dim())
        assert y.shape[1] == self.out_channels, f"Expected {self.out_channels} channels, got {y.shape[1]}"
        assert y.shape[2] == x.shape[2] and y.shape[3] == x.shape[3], \
            f"Expected spatial dimensions {(x.shape[2], x.shape[3])}, got {(y.shape[2], y.shape[3])}"
        return y


class AdaptiveGroupNorm(nn.Module):
    """model.py:304-316: GroupNorm(32, C) (affine) followed by a second learned per-channel affine."""

    def __init__(self, num_channels, num_groups=32):
        super().__init__()
        self.num_channels = num_channels
        self.num_groups = num_groups
        self.weight = nn.Parameter(torch.ones(1, num_channels, 1, 1, 1))
        self.bias = nn.Parameter(torch.zeros(1, num_channels, 1, 1, 1))
        self.group_norm = nn.GroupNorm(num_groups, num_channels)

    def _affine(self):
        return (_f32c(self.group_norm.weight), _f32c(self.group_norm.bias), _f32c(self.weight).view(-1),
                _f32c(self.bias).view(-1))

    def _forward_autograd(self, x, stats=None):
        """Differentiable form (row f-2): GroupNorm forward / backward on libmpb200, the second affine as a torch expression."""
        gn = self.group_norm
        return ops.GroupNormFunction.apply(x, self.num_groups, gn.weight, gn.bias, gn.eps, stats) * self.weight + self.bias

    def forward(self, x):
        if _wants_grad(self, x):
            _require_inference(self, x.detach())
            return _public(self._forward_autograd(x.float()))
        _require_inference(self, x)
        a = ops.from_nchw(_as_f32_cuda(x), f32=True, split=False)
        g, b, g2, b2 = self._affine()
        out = ops.group_norm_act(a, self.num_groups, None, g, b, g2, b2, act=ACT_NONE, f32=True, split=False)
        return ops.to_nchw(out, x.dim())


class ResBlock3D_Adaptive(nn.Module, _Packed):
    """model.py:369-408."""

    def __init__(self, in_channels, out_channels, upsample=False, scale_factors=(1, 1, 1)):
        super().__init__()
        self.upsample = upsample
        self.scale_factors = scale_factors
        self.conv1 = nn.Conv3d(in_channels, out_channels, 3, padding=1)
        self.conv2 = nn.Conv3d(out_channels, out_channels, 3, padding=1)
        self.norm1 = AdaptiveGroupNorm(out_channels)
        self.norm2 = AdaptiveGroupNorm(out_channels)
        if in_channels != out_channels:
            self.residual_conv = nn.Conv3d(in_channels, out_channels, 1)
        else:
            self.residual_conv = nn.Identity()

    def _build_plan(self):
        dev = self.conv1.weight.device
        P = {"c1": ops.pack_conv(self.conv1.weight, self.conv1.bias, dev),
             "c2": ops.pack_conv(self.conv2.weight, self.conv2.bias, dev),
             "n1": self.norm1._affine(), "n2": self.norm2._affine(), "rc": None}
        if isinstance(self.residual_conv, nn.Conv3d):
            P["rc"] = ops.pack_conv(self.residual_conv.weight, self.residual_conv.bias, dev)
        return P

    def _forward_cl(self, x: Act, f32: bool = True, split: bool = True, _module_call: bool = False) -> Act:
        """x: split (+ f32 when the residual is the identity and full precision is wanted)."""
        if self.upsample and not _module_call:
            raise NotImplementedError("ResBlock3D_Adaptive(upsample=True) inside a fused pipeline: the reference has no such call "
                                      "site (the stand-alone module forward() supports it)")
        P = self._plan()
        G = self.norm1.num_groups
        h, st = ops.conv(x, P["c1"], f32=True, stats_groups=G)
        h = ops.group_norm_act(h, G, st, *P["n1"], act=ACT_RELU, split=True)
        h, st = ops.conv(h, P["c2"], f32=True, stats_groups=G)
        res = x if P["rc"] is None else ops.conv(x, P["rc"], f32=True)[0]
        return ops.group_norm_act(h, G, st, *P["n2"], res=res, act=ACT_RELU, f32=f32, split=split)

    def _tail(self, y):
        """`upsample=True` (model.py:405-406; no call site in the reference): the trailing trilinear resize, through ATen."""
        if self.upsample:
            y = F.interpolate(y, scale_factor=self.scale_factors, mode="trilinear", align_corners=False)
        return y

    def _forward_autograd(self, x):
        """Differentiable form (row f-2)."""
        conv = ops.conv_train
        t, st = conv(x, self.conv1.weight, self.conv1.bias, stats_groups=self.norm1.num_groups)
        h = torch.relu(self.norm1._forward_autograd(t, st))
        t, st = conv(h, self.conv2.weight, self.conv2.bias, stats_groups=self.norm2.num_groups)
        h = self.norm2._forward_autograd(t, st)
        res = conv(x, self.residual_conv.weight, self.residual_conv.bias) if isinstance(self.residual_conv, nn.Conv3d) else x
        return torch.relu(h + res)

    def forward(self, x):
        if _wants_grad(self, x):
            _require_inference(self, x.detach())
            return self._tail(_public(self._forward_autograd(x.float())))
        _require_inference(self, x)
        a = ops.from_nchw(_as_f32_cuda(x), f32=True, split=True)
        return self._tail(ops.to_nchw(self._forward_cl(a, f32=True, split=False, _module_call=True), 5))


class FlowField(nn.Module, _Packed):
    """model.py:415-471: 512-vector -> (B,3,16,16,16) flow in [0,1)."""

    def __init__(self):
        super().__init__()
        self.conv1x1 = nn.Conv2d(512, 2048, kernel_size=1)
        self.reshape_layer = lambda x: x.view(-1, 512, 4, *x.shape[2:])
        self.resblock1 = ResBlock3D_Adaptive(in_channels=512, out_channels=256)
        self.upsample1 = nn.Upsample(scale_factor=(2, 2, 2))
        self.resblock2 = ResBlock3D_Adaptive(in_channels=256, out_channels=128)
        self.upsample2 = nn.Upsample(scale_factor=(2, 2, 2))
        self.resblock3 = ResBlock3D_Adaptive(in_channels=128, out_channels=64)
        self.upsample3 = nn.Upsample(scale_factor=(1, 2, 2))
        self.resblock4 = ResBlock3D_Adaptive(in_channels=64, out_channels=32)
        self.upsample4 = nn.Upsample(scale_factor=(1, 2, 2))
        self.conv3x3x3 = nn.Conv3d(32, 3, kernel_size=3, padding=1)
        self.gn = nn.GroupNorm(1, 3)
        self.tanh = nn.Tanh()

    def _build_plan(self):
        dev = self.conv1x1.weight.device
        # output channel c*4+d of the 1x1 conv becomes (c, d) of the (512,4,1,1) volume (model.py:424): permute the
        # rows to d*512+c so the GEMM result is already the channels-last [B, 4, 1, 1, 512] volume.
        w = self.conv1x1.weight.detach().view(512, 4, 512, 1, 1).permute(1, 0, 2, 3, 4).reshape(2048, 512, 1, 1)
        b = self.conv1x1.bias.detach().view(512, 4).t().reshape(2048)
        return {"c0": ops.pack_conv(w, b, dev),
                "c5": ops.pack_conv(self.conv3x3x3.weight, self.conv3x3x3.bias, dev),
                "gn": (_f32c(self.gn.weight), _f32c(self.gn.bias))}

    def _forward_cl(self, s: torch.Tensor) -> torch.Tensor:
        """s [B,512] fp32 -> em channels-last [B,16,16,16,3] fp32."""
        P = self._plan()
        B = s.shape[0]
        a = Act((B, 1, 1, 1, 512), f32=s.contiguous().view(B, 1, 1, 1, 512))
        h, _ = ops.conv(a, P["c0"], f32=True)
        x = Act((B, 4, 1, 1, 512), f32=h.f32.view(B, 4, 1, 1, 512))
        ops.ensure_split(x)
        for blk, sc in ((self.resblock1, (2, 2, 2)), (self.resblock2, (2, 2, 2)), (self.resblock3, (1, 2, 2)),
                        (self.resblock4, (1, 2, 2))):
            y = blk._forward_cl(x, f32=True, split=False)
            x = ops.upsample_nearest(y, sc, f32=False, split=True)
        h, st = ops.conv(x, P["c5"], f32=True, stats_groups=1)
        out = ops.group_norm_act(h, 1, st, *P["gn"], act=ACT_RELU_TANH, f32=True, split=False)
        return out.f32

    def _forward_autograd(self, zs):
        """Differentiable form (row f-2): zs [B,512,1,1] -> (B,3,16,16,16)."""
        x = ops.conv_train(zs, self.conv1x1.weight, self.conv1x1.bias)
        x = x.view(-1, 512, 4, *x.shape[2:])
        for blk, sf in ((self.resblock1, (2, 2, 2)), (self.resblock2, (2, 2, 2)), (self.resblock3, (1, 2, 2)),
                        (self.resblock4, (1, 2, 2))):
            x = ops.UpsampleNearestFunction.apply(blk._forward_autograd(x), sf[0])
        x = ops.conv_train(x, self.conv3x3x3.weight, self.conv3x3x3.bias)
        x = ops.GroupNormFunction.apply(x, 1, self.gn.weight, self.gn.bias, self.gn.eps)
        return torch.tanh(torch.relu(x))

    def forward(self, zs, adaptive_gamma, adaptive_beta):
        if _wants_grad(self, zs):
            _require_inference(self, zs.detach())
            x = _public(self._forward_autograd(zs.float().reshape(zs.shape[0], 512, 1, 1)))
            assert x.shape[1] == 3, f"Expected 3 channels after conv3x3x3, got {x.shape[1]}"
            return x
        _require_inference(self, zs)
        em = self._forward_cl(_as_f32_cuda(zs).reshape(zs.shape[0], 512))
        x = em.permute(0, 4, 1, 2, 3).contiguous()
        assert x.shape[1] == 3, f"Expected 3 channels after conv3x3x3, got {x.shape[1]}"
        return x


class ResBlock3D(nn.Module, _Packed):
    """model.py:500-528."""

    def __init__(self, in_channels, out_channels, upsample=False, scale_factors=(1, 1, 1)):
        super().__init__()
        self.upsample = upsample
        self.scale_factors = scale_factors
        self.conv1 = nn.Conv3d(in_channels, out_channels, kernel_size=3, padding=1)
        self.gn1 = nn.GroupNorm(num_groups=32, num_channels=out_channels)
        self.conv2 = nn.Conv3d(out_channels, out_channels, kernel_size=3, padding=1)
        self.gn2 = nn.GroupNorm(num_groups=32, num_channels=out_channels)
        self.shortcut = nn.Conv3d(in_channels, out_channels, kernel_size=1) if in_channels != out_channels \
            else nn.Identity()

    def _build_plan(self):
        dev = self.conv1.weight.device
        P = {"c1": ops.pack_conv(self.conv1.weight, self.conv1.bias, dev),
             "c2": ops.pack_conv(self.conv2.weight, self.conv2.bias, dev),
             "g1": (_f32c(self.gn1.weight), _f32c(self.gn1.bias)),
             "g2": (_f32c(self.gn2.weight), _f32c(self.gn2.bias)), "sc": None}
        if isinstance(self.shortcut, nn.Conv3d):
            P["sc"] = ops.pack_conv(self.shortcut.weight, self.shortcut.bias, dev)
        return P

    def _forward_cl(self, x: Act, f32: bool = True, split: bool = False, _module_call: bool = False) -> Act:
        if self.upsample and not _module_call:
            raise NotImplementedError("ResBlock3D(upsample=True) inside a fused pipeline: the reference has no such call site "
                                      "(the stand-alone module forward() supports it)")
        P = self._plan()
        idt = x if P["sc"] is None else ops.conv(x, P["sc"], f32=True)[0]
        h, st = ops.conv(x, P["c1"], f32=True, stats_groups=32)
        h = ops.group_norm_act(h, 32, st, *P["g1"], act=ACT_RELU, split=True)
        h, st = ops.conv(h, P["c2"], f32=True, stats_groups=32)
        return ops.group_norm_act(h, 32, st, *P["g2"], res=idt, act=ACT_RELU, f32=f32, split=split)

    def _forward_autograd(self, x):
        """Differentiable form (row f-2): the same operators as autograd Functions whose forward AND backward run on libmpb200
        (ops.ConvFunction: tcgen05 forward / data gradient, tensor-core weight gradient; ops.GroupNormFunction)."""
        conv, gn = ops.conv_train, ops.GroupNormFunction.apply
        idt = conv(x, self.shortcut.weight, self.shortcut.bias) if isinstance(self.shortcut, nn.Conv3d) else x
        t, st = conv(x, self.conv1.weight, self.conv1.bias, stats_groups=32)      # statistics from the convolution's epilogue
        h = torch.relu(gn(t, 32, self.gn1.weight, self.gn1.bias, self.gn1.eps, st))
        t, st = conv(h, self.conv2.weight, self.conv2.bias, stats_groups=32)
        h = gn(t, 32, self.gn2.weight, self.gn2.bias, self.gn2.eps, st)
        return torch.relu(h + idt)

    def _tail(self, y):
        """`upsample=True` (model.py:525-526; no call site in the reference): the trailing trilinear resize, through ATen."""
        if self.upsample:
            y = F.interpolate(y, scale_factor=self.scale_factors, mode="trilinear", align_corners=False)
        return y

    def forward(self, x):
        if _wants_grad(self, x):
            _require_inference(self, x.detach())         # (device check only)
            return self._tail(_public(self._forward_autograd(x.float())))
        _require_inference(self, x)
        a = ops.from_nchw(_as_f32_cuda(x), f32=True, split=True)
        return self._tail(ops.to_nchw(self._forward_cl(a, f32=True, split=False, _module_call=True), 5))


class G3d(nn.Module):
    """model.py:571-597."""

    def __init__(self, in_channels):
        super().__init__()
        self.downsampling = nn.Sequential(
            ResBlock3D(in_channels, 96), nn.AvgPool3d(kernel_size=2, stride=2),
            ResBlock3D(96, 192), nn.AvgPool3d(kernel_size=2, stride=2),
            ResBlock3D(192, 384), nn.AvgPool3d(kernel_size=2, stride=2),
            ResBlock3D(384, 768))
        self.upsampling = nn.Sequential(
            ResBlock3D(768, 384), nn.Upsample(scale_factor=2, mode='trilinear', align_corners=True),
            ResBlock3D(384, 192), nn.Upsample(scale_factor=2, mode='trilinear', align_corners=True),
            ResBlock3D(192, 96), nn.Upsample(scale_factor=2, mode='trilinear', align_corners=True))
        self.final_conv = nn.Conv3d(96, 96, kernel_size=3, padding=1)

    def _final_pack(self):
        sig = (self.final_conv.weight._version, self.final_conv.bias._version, str(self.final_conv.weight.device),
               self.final_conv.weight.data_ptr())
        c = self.__dict__.get("_mp_final")
        if c is None or c[0] != sig:
            c = (sig, ops.pack_conv(self.final_conv.weight, self.final_conv.bias))
            self.__dict__["_mp_final"] = c
        return c[1]

    def _forward_cl(self, x: Act) -> Act:
        """x: f32 + split, channels-last [N,16,64,64,96] -> f32 channels-last."""
        for i in (0, 2, 4, 6):
            y = self.downsampling[i]._forward_cl(x, f32=True, split=(i == 6))
            x = ops.avgpool2(y, 2, f32=False, split=True) if i < 6 else y
        for i in (0, 2, 4):
            y = self.upsampling[i]._forward_cl(x, f32=True, split=False)
            x = ops.upsample2x_linear(y, 2, f32=False, split=True)
        out, _ = ops.conv(x, self._final_pack(), f32=True)
        return out

    def _forward_autograd(self, x):
        """Differentiable form (row f-2): residual blocks, the final convolution, the 2x average pools and the trilinear upsamples
        between them through the libmpb200 Functions of ops.py (CUDA forward and backward)."""
        for i, m in enumerate(self.downsampling):
            x = m._forward_autograd(x) if i % 2 == 0 else ops.AvgPool2Function.apply(x, 2)
        for i, m in enumerate(self.upsampling):
            x = m._forward_autograd(x) if i % 2 == 0 else ops.UpsampleLinear2xFunction.apply(x)
        return ops.ConvFunction.apply(x, self.final_conv.weight, self.final_conv.bias)

    def forward(self, x):
        if _wants_grad(self, x):
            _require_inference(self, x.detach())
            return _public(self._forward_autograd(x.float()))
        _require_inference(self, x)
        a = ops.from_nchw(_as_f32_cuda(x), f32=True, split=True)
        return ops.to_nchw(self._forward_cl(a), 5)


class ResBlock2D(nn.Module, _Packed):
    """model.py:600-640 (BatchNorm2d: eval-mode running statistics are folded into the conv weights)."""

    def __init__(self, in_channels, out_channels, downsample=False):
        super().__init__()
        self.downsample = downsample
        self.conv1 = nn.Conv2d(in_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.bn1 = nn.BatchNorm2d(out_channels)
        self.conv2 = nn.Conv2d(out_channels, out_channels, kernel_size=3, stride=1, padding=1)
        self.bn2 = nn.BatchNorm2d(out_channels)
        if self.downsample:
            self.downsample_conv = nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=2)
            self.downsample_bn = nn.BatchNorm2d(out_channels)
        if in_channels != out_channels:
            self.shortcut = nn.Sequential(nn.Conv2d(in_channels, out_channels, kernel_size=1, stride=1),
                                          nn.BatchNorm2d(out_channels))
        else:
            self.shortcut = nn.Identity()

    @staticmethod
    def _bn(bn: nn.BatchNorm2d) -> dict:
        return {"weight": bn.weight.detach(), "bias": bn.bias.detach(), "running_mean": bn.running_mean,
                "running_var": bn.running_var}

    def _build_plan(self):
        if self.training:
            raise NotImplementedError("ResBlock2D: train-mode BatchNorm (batch statistics) is not implemented "
                                      "(SURVEY.md 8f-2); call .eval()")
        if self.downsample:
            raise NotImplementedError("ResBlock2D(downsample=True): no call site in the reference, and its own forward fails there "
                                      "(model.py:632-637 adds a stride-2 identity to a full-resolution tensor)")
        sc = None
        if isinstance(self.shortcut, nn.Sequential):
            # Conv2d(1x1) + BatchNorm shortcut: folded, then fused into conv2's accumulator as extra K columns, so the
            # shortcut tensor never exists in HBM (out = relu(W2 * t + Ws . x + b2 + bs))
            sc = ops.fold_bn(self.shortcut[0].weight.detach(), self.shortcut[0].bias.detach(),
                             self._bn(self.shortcut[1]), self.shortcut[1].eps)
        # identity-shortcut blocks inside G2d run the fp16 + FP8 cross-term convolution (`G2d` sets `_mp_q8`): two
        # pass-units instead of three at the same end-to-end accuracy (tests/precision_study.py)
        q8 = getattr(self, "_mp_q8", False) and _Q8_ENABLED
        if sc is not None and q8:
            # up-blocks (round 2): the same format with the fused shortcut; channel counts must be 64-multiples on both sources
            q8 = self.conv1.in_channels % 64 == 0 and self.conv1.out_channels % 128 == 0
        prec = ops.PREC_F16_Q8 if q8 else ops.PREC_SPLIT_BF16
        c1, c2 = self._pack_pair(prec, sc)
        P = {"c1": c1, "c2": c2, "fused_sc": sc is not None, "q8": q8}
        if q8 and sc is not None:       # kept beside the FP8 plans: the fall-back of G2d's range check and stand-alone calls
            P["c1_split"], P["c2_split"] = self._pack_pair(ops.PREC_SPLIT_BF16, sc)
        return P

    def _pack_pair(self, prec, sc=None):
        dev = self.conv1.weight.device
        w1, b1 = ops.fold_bn(self.conv1.weight.detach(), self.conv1.bias.detach(), self._bn(self.bn1), self.bn1.eps)
        w2, b2 = ops.fold_bn(self.conv2.weight.detach(), self.conv2.bias.detach(), self._bn(self.bn2), self.bn2.eps)
        return ops.pack_conv(w1, b1, dev, prec=prec), ops.pack_conv(w2, b2, dev, shortcut=sc, prec=prec)

    def _forward_cl(self, x: Act, f32: bool = False, split: bool = True, stats_groups: int = 0, hq_out: bool = False,
                    q8_scales: Tuple[float, float] = (1.0, 1.0)):
        """x: split (or fp16 + FP8 planes for a `_mp_q8` block).  Returns (Act, stats or None).  `q8_scales`: the
        per-tensor scales of the FP8 byte planes this block writes (its inner activation, its output)."""
        P = self._plan()
        if P["q8"] and x.q8 is not None:
            if P["fused_sc"]:    # up-block: x (the upsampled tensor) and t share one byte-plane scale (one cross accumulator)
                t, _ = ops.conv(x, P["c1"], act=ACT_RELU, f32=False, hq=True, out_q8_scale=x.q8_scale)
                return ops.conv(t, P["c2"], src2=x, act=ACT_RELU, f32=f32, split=split)
            t, _ = ops.conv(x, P["c1"], act=ACT_RELU, f32=False, hq=True, out_q8_scale=q8_scales[0])
            return ops.conv(t, P["c2"], res=x, act=ACT_RELU, f32=f32, split=split and not hq_out, hq=hq_out,
                            out_q8_scale=q8_scales[1])
        if P["q8"]:          # a G2d block called on its own with split planes: three-pass packs, built on first use
            if "c1_split" not in P:
                P["c1_split"], P["c2_split"] = self._pack_pair(ops.PREC_SPLIT_BF16)
            t, _ = ops.conv(x, P["c1_split"], act=ACT_RELU, f32=False, split=True)
            if P["fused_sc"]:
                return ops.conv(t, P["c2_split"], src2=x, act=ACT_RELU, f32=f32, split=split, stats_groups=stats_groups)
            return ops.conv(t, P["c2_split"], res=x, act=ACT_RELU, f32=f32, split=split, stats_groups=stats_groups)
        t, _ = ops.conv(x, P["c1"], act=ACT_RELU, f32=False, split=True)
        if P["fused_sc"]:
            return ops.conv(t, P["c2"], src2=x, act=ACT_RELU, f32=f32, split=split, stats_groups=stats_groups)
        return ops.conv(t, P["c2"], res=x, act=ACT_RELU, f32=f32, split=split, stats_groups=stats_groups)

    def _forward_autograd(self, x):
        """Differentiable / train-mode form (row f-2): convolutions and BatchNorm (batch statistics in train mode, running
        statistics updated like ATen does) through the libmpb200 Functions of ops.py."""
        if self.downsample:
            raise NotImplementedError("ResBlock2D(downsample=True): no call site in the reference, and its own forward fails there "
                                      "(model.py:632-637 adds a stride-2 identity to a full-resolution tensor)")
        cb = ops.conv_bn_train            # (train mode: the batch statistics come out of the convolution's epilogue)
        out = torch.relu(cb(x, self.conv1, self.bn1))
        out = cb(out, self.conv2, self.bn2)
        if isinstance(self.shortcut, nn.Sequential):
            x = cb(x, self.shortcut[0], self.shortcut[1])
        return torch.relu(out + x)

    def forward(self, x):
        if self.training or _wants_grad(self, x):
            _require_inference(self, x.detach())
            return _public(self._forward_autograd(x.float()))
        _require_inference(self, x)
        a = ops.from_nchw(_as_f32_cuda(x), f32=False, split=True)
        out, _ = self._forward_cl(a, f32=True, split=False)
        return ops.to_nchw(out, 4)


class G2d(nn.Module, _Packed):
    """model.py:715-763."""

    def __init__(self, in_channels):
        super().__init__()
        self.reshape = nn.Conv2d(96, 1536, kernel_size=1)
        self.conv1x1 = nn.Conv2d(1536, 512, kernel_size=1)
        self.res_blocks = nn.Sequential(*[ResBlock2D(512, 512) for _ in range(8)])
        for blk in self.res_blocks:
            blk._mp_q8 = True     # identity res-blocks: fp16 + FP8 cross-term convolutions (see ResBlock2D._build_plan)
        self.upsample1 = nn.Sequential(nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True),
                                       ResBlock2D(512, 256))
        self.upsample2 = nn.Sequential(nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True),
                                       ResBlock2D(256, 128))
        self.upsample3 = nn.Sequential(nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True),
                                       ResBlock2D(128, 64))
        self.final_conv = nn.Sequential(nn.GroupNorm(num_groups=32, num_channels=64), nn.ReLU(inplace=True),
                                        nn.Conv2d(64, 3, kernel_size=3, padding=1), nn.Sigmoid())
        if _Q8_UP:                # up-blocks 512->256 @128^2 and 256->128 @256^2 on the FP8 cross-term format as well
            self.upsample1[1]._mp_q8 = True
            self.upsample2[1]._mp_q8 = True

    def _sig_extra(self):
        return None

    def _build_plan(self):
        dev = self.reshape.weight.device
        # reshape (96->1536) and conv1x1 (1536->512) have no non-linearity between them (model.py:756-757):
        # fold them into one 96->512 1x1 conv in float64.
        w1 = self.reshape.weight.detach().double().view(1536, 96)
        b1 = self.reshape.bias.detach().double()
        w2 = self.conv1x1.weight.detach().double().view(512, 1536)
        b2 = self.conv1x1.bias.detach().double()
        w = (w2 @ w1).view(512, 96, 1, 1)
        b = w2 @ b1 + b2
        return {"in": ops.pack_conv(w, b, dev),
                "gn": (_f32c(self.final_conv[0].weight), _f32c(self.final_conv[0].bias)),
                # GroupNorm -> ReLU -> 64->3 3x3 -> Sigmoid runs as one fused pass; its weights travel as kernel parameters
                "head_w": self.final_conv[2].weight.detach().float().cpu().contiguous(),
                "head_b": self.final_conv[2].bias.detach().float().cpu().contiguous()}

    # fp16 + FP8 cross-term mode: safe range of a tensor's largest magnitude.  Above Q8_AMAX_LIMIT the fp16 main plane
    # itself is close to overflow (65504): the chain then runs the three-pass split-bf16 plans instead.
    Q8_AMAX_LIMIT = 16384.0

    def _res_chain(self, h: Act, scales, collect=None) -> Act:
        """The 8 identity res-blocks on F16_Q8 planes; `scales[1 + 2 i]`, `scales[2 + 2 i]` = byte-plane scales of block
        i's inner activation and output.  `collect` (calibration) receives every F16_Q8 tensor of the chain."""
        n = len(self.res_blocks)
        for i, blk in enumerate(self.res_blocks):
            P = blk._plan()
            t, _ = ops.conv(h, P["c1"], act=ACT_RELU, f32=False, hq=True, out_q8_scale=scales[1 + 2 * i])
            last = i + 1 == n
            h2, _ = ops.conv(t, P["c2"], res=h, act=ACT_RELU, f32=False, split=last, hq=not last,
                             out_q8_scale=scales[2 + 2 * i])
            if collect is not None:
                collect.append(t)
                if not last:
                    collect.append(h2)
            h = h2
        return h

    def _up_q8(self, i: int) -> bool:
        blk = (self.upsample1, self.upsample2, self.upsample3)[i][1]
        return bool(blk._plan()["q8"] and blk._plan()["fused_sc"])

    def _q8_calibrate(self, x: Act, P) -> None:
        """One pass of the res-block chain (and of the up-blocks that use the format) with unit scales to measure every
        F16_Q8 tensor's largest magnitude, then per-tensor power-of-two scales for the FP8 byte planes
        (ops.q8_scale_for); tensors near the fp16 range switch the stage to the split-bf16 plans.  Runs once per plan,
        outside CUDA-graph capture (one host sync)."""
        ones = [1.0] * (1 + 2 * len(self.res_blocks))
        h, _ = ops.conv(x, P["in"], f32=False, split=False, hq=True)
        seen = [h]
        h = self._res_chain(h, ones, collect=seen)
        up_seen = []
        for i, up in enumerate((self.upsample1, self.upsample2)):
            if not self._up_q8(i):
                break
            Pb = up[1]._plan()
            u = ops.upsample2x_bilinear_hq(h, 1.0)
            t, _ = ops.conv(u, Pb["c1"], act=ACT_RELU, f32=False, hq=True)
            h, _ = ops.conv(t, Pb["c2"], src2=u, act=ACT_RELU, f32=False, split=True)
            up_seen += [u, t]
        amax = torch.stack([t.h16.float().abs().amax() for t in seen + up_seen]).cpu().tolist()
        a_chain, a_up = amax[:len(seen)], amax[len(seen):]
        ok = all(a == a and a < self.Q8_AMAX_LIMIT for a in a_chain)       # NaN / inf / near-overflow -> fall back
        P["q8_amax"] = amax
        # (the last block writes split-bf16 planes: its output has no byte plane, the trailing 1.0 is never used)
        P["q8_scales"] = [ops.q8_scale_for(a) for a in a_chain] + [1.0] if ok else None
        P["q8_ok"] = ok
        # up-blocks: the upsampled tensor and the block's inner activation share ONE scale (fused shortcut: one accumulator)
        P["q8_up"] = []
        for j in range(0, len(a_up), 2):
            good = ok and all(a == a and a < self.Q8_AMAX_LIMIT for a in a_up[j:j + 2])
            P["q8_up"].append(min(ops.q8_scale_for(a_up[j]), ops.q8_scale_for(a_up[j + 1])) if good else None)

    def _forward_cl(self, x: Act) -> torch.Tensor:
        """x: split channels-last [N,1,64,64,96] -> RGB NCHW fp32 [N,3,512,512]."""
        P = self._plan()
        q8 = _Q8_ENABLED and all(blk._plan()["q8"] for blk in self.res_blocks)
        scales = None
        if q8:
            if "q8_ok" not in P and not _capturing(x):
                self._q8_calibrate(x, P)
            q8 = P.get("q8_ok", True)
            scales = P.get("q8_scales") or [1.0] * (1 + 2 * len(self.res_blocks))
        if q8:
            h, _ = ops.conv(x, P["in"], f32=False, split=False, hq=True, out_q8_scale=scales[0])
            h = self._res_chain(h, scales)
        else:
            h, _ = ops.conv(x, P["in"], f32=False, split=True)
            for blk in self.res_blocks:
                h, _ = blk._forward_cl(h)
        st = None
        up_scales = P.get("q8_up", []) if q8 else []
        for i, up in enumerate((self.upsample1, self.upsample2, self.upsample3)):
            last = i == 2
            s_up = up_scales[i] if i < len(up_scales) else None
            if s_up is not None and self._up_q8(i):
                u = ops.upsample2x_bilinear_hq(h, s_up)
            else:
                u = ops.upsample2x_linear(h, 1, f32=False, split=True)
            h, st = up[1]._forward_cl(u, f32=last, split=not last, stats_groups=32 if last else 0)
        ab = ops.gn_finalize(st, h.shape, 32, *P["gn"])
        return ops.gn_relu_conv3x3_head(h, ab, P["head_w"], P["head_b"], ACT_SIGMOID)

    def _forward_autograd(self, x):
        """Differentiable / train-mode form (row f-2).  `reshape` and `conv1x1` have no non-linearity between them (model.py:756-757):
        their product is formed as a torch expression on the two weights (autograd carries the gradient to both) and ONE 96 -> 512
        convolution runs; the bilinear x2 upsamples are `ops.UpsampleLinear2xFunction` (CUDA forward and backward)."""
        conv = ops.conv_train
        w2 = self.conv1x1.weight.view(512, 1536)
        w = (w2 @ self.reshape.weight.view(1536, 96)).view(512, 96, 1, 1)
        x = conv(x, w, w2 @ self.reshape.bias + self.conv1x1.bias)
        for blk in self.res_blocks:
            x = blk._forward_autograd(x)
        for up in (self.upsample1, self.upsample2, self.upsample3):
            x = up[1]._forward_autograd(ops.UpsampleLinear2xFunction.apply(x))
        gn = self.final_conv[0]
        x = torch.relu(ops.GroupNormFunction.apply(x, gn.num_groups, gn.weight, gn.bias, gn.eps))
        return torch.sigmoid(conv(x, self.final_conv[2].weight, self.final_conv[2].bias))

    def forward(self, x):
        if self.training or _wants_grad(self, x):
            _require_inference(self, x.detach())
            return _public(self._forward_autograd(x.float()))
        _require_inference(self, x)
        a = ops.from_nchw(_as_f32_cuda(x), f32=False, split=True)
        return self._forward_cl(a)


# ----------------------------------------------------------------------------------------------------- Eapp
class Eapp(nn.Module, _Packed):
    """model.py:206-299 (including the double assignment of `resblock3D_96_2`, model.py:218,225)."""

    def __init__(self):
        super().__init__()
        self.conv = nn.Conv2d(3, 64, 7, stride=1, padding=3)
        self.resblock_128 = ResBlock_Custom(dimension=2, in_channels=64, out_channels=128)
        self.resblock_256 = ResBlock_Custom(dimension=2, in_channels=128, out_channels=256)
        self.resblock_512 = ResBlock_Custom(dimension=2, in_channels=256, out_channels=512)
        self.resblock3D_96 = ResBlock3D_Adaptive(in_channels=96, out_channels=96)
        self.resblock3D_96_2 = ResBlock3D_Adaptive(in_channels=96, out_channels=96)
        self.resblock3D_96_1 = ResBlock3D_Adaptive(in_channels=96, out_channels=96)
        self.resblock3D_96_1_2 = ResBlock3D_Adaptive(in_channels=96, out_channels=96)
        self.resblock3D_96_2 = ResBlock3D_Adaptive(in_channels=96, out_channels=96)
        self.resblock3D_96_2_2 = ResBlock3D_Adaptive(in_channels=96, out_channels=96)
        self.conv_1 = nn.Conv2d(in_channels=512, out_channels=1536, kernel_size=1, stride=1, padding=0)
        self.avgpool = nn.AvgPool2d(kernel_size=2, stride=2, padding=0)
        self.custom_resnet50 = CustomResNet50()
        self.fc = torch.nn.Linear(2048, COMPRESS_DIM)
        self.tf32_descriptor = True   # only used by the stock cuDNN backend of the ResNet-50 descriptor branch

    def _build_plan(self):
        dev = self.conv.weight.device
        # RGB stem on the tensor-core kernel: 3 input channels zero-padded to 16
        return {"stem": ops.pack_conv(self.conv.weight, self.conv.bias, dev, cin_pad=16),
                "c1": ops.pack_conv(self.conv_1.weight, self.conv_1.bias, dev)}

    def _volume_cl(self, x: torch.Tensor) -> Act:
        """x NCHW fp32 [B,3,512,512] -> vs channels-last [B,16,64,64,96] (f32 + split)."""
        P = self._plan()
        from .emtn_cuda import rgb16
        a = rgb16(x)
        out, st = ops.conv(a, P["stem"], f32=True, split=True, stats_groups=32)
        for blk in (self.resblock_128, self.resblock_256, self.resblock_512):
            y = blk._forward_cl(out, st)
            out = ops.avgpool2(y, 1, f32=True, split=True)
            st = None
        h = ops.group_norm_act(out, 32, None, act=ACT_RELU, split=True)
        h, _ = ops.conv(h, P["c1"], f32=True)
        B, _, Hh, Ww, _ = h.shape
        # view(B, 96, 16, H, W) (model.py:271): channel c*16+d -> (c, d); go through NCHW to re-tile as NDHWC
        vol = ops.to_nchw(h, 4).view(B, 96, 16, Hh, Ww)
        vs = ops.from_nchw(vol, f32=True, split=True)
        for blk in (self.resblock3D_96, self.resblock3D_96_2, self.resblock3D_96_1, self.resblock3D_96_1_2,
                    self.resblock3D_96_2, self.resblock3D_96_2_2):
            vs = blk._forward_cl(vs, f32=True, split=True)
        return vs

    def _descriptor(self, x: torch.Tensor) -> torch.Tensor:
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=bool(self.tf32_descriptor)):
            es = self.custom_resnet50(x)
        return self.fc(torch.flatten(es, start_dim=1))

    def _forward_autograd(self, x):
        """Differentiable / train-mode form (row f-2) -> (vs, es)."""
        conv, gn = ops.conv_train, ops.GroupNormFunction.apply
        out = conv(x, self.conv.weight, self.conv.bias)
        for blk in (self.resblock_128, self.resblock_256, self.resblock_512):
            out = ops.AvgPool2Function.apply(blk._forward_autograd(out), 1)
        out = conv(torch.relu(gn(out, 32, None, None, 1e-5)), self.conv_1.weight, self.conv_1.bias)
        vs = out.view(out.size(0), 96, 16, *out.shape[2:])
        for blk in (self.resblock3D_96, self.resblock3D_96_2, self.resblock3D_96_1, self.resblock3D_96_1_2,
                    self.resblock3D_96_2, self.resblock3D_96_2_2):
            vs = blk._forward_autograd(vs)
        es = self.custom_resnet50._forward_autograd(x)
        return vs, self.fc(torch.flatten(es, start_dim=1))

    def forward(self, x):
        if self.training or _wants_grad(self, x):
            _require_inference(self, x.detach())
            return _public(self._forward_autograd(x.float()))
        _require_inference(self, x)
        x = _as_f32_cuda(x)
        vs = ops.to_nchw(self._volume_cl(x), 5)
        return vs, self._descriptor(x)


# ----------------------------------------------------------------------------------------------------- warps
def compute_rotation_matrix(rotation):
    """model.py:811-856: degrees -> R = Rx(a) @ (Ry(b) @ Rz(c))."""
    r = rotation * (torch.pi / 180.0)
    ca, sa = torch.cos(r[:, 0]), torch.sin(r[:, 0])
    cb, sb = torch.cos(r[:, 1]), torch.sin(r[:, 1])
    cg, sg = torch.cos(r[:, 2]), torch.sin(r[:, 2])
    z, o = torch.zeros_like(ca), torch.ones_like(ca)
    Rx = torch.stack((torch.stack((o, z, z), 1), torch.stack((z, ca, -sa), 1), torch.stack((z, sa, ca), 1)), 1)
    Ry = torch.stack((torch.stack((cb, z, sb), 1), torch.stack((z, o, z), 1), torch.stack((-sb, z, cb), 1)), 1)
    Rz = torch.stack((torch.stack((cg, -sg, z), 1), torch.stack((sg, cg, z), 1), torch.stack((z, z, o), 1)), 1)
    return torch.matmul(Rx, torch.matmul(Ry, Rz))


def _affine_3x4(rotation, translation, invert: bool) -> torch.Tensor:
    """First three rows of [R|t] or of its inverse (model.py:790-803) as a contiguous [B,3,4] fp32 tensor."""
    B = rotation.shape[0]
    R = compute_rotation_matrix(rotation.float())
    t = translation.float().unsqueeze(-1)
    if invert:
        # inverse of the rigid transform [R|t; 0 0 0 1] in closed form ([R^T | -R^T t]); the reference calls
        # torch.inverse on the same matrix (model.py:803) -- identical up to fp32 rounding, but free of the host
        # synchronisation of the LU path, so the step can be captured in a CUDA graph
        Rt = R.transpose(1, 2)
        return torch.cat((Rt, -torch.matmul(Rt, t)), dim=2).contiguous()
    return torch.cat((R, t), dim=2).contiguous()


def compute_rt_warp(rotation, translation, invert=False, grid_size=64):
    """model.py:777-809 -> (B, 3, G, G, G)."""
    if torch.is_grad_enabled() and (rotation.requires_grad or translation.requires_grad):
        # differentiable form (row f-2): the 3-channel rigid grid through ATen's autograd, as inside the warp generators
        theta = _affine_3x4(rotation, translation, invert)
        grid = F.affine_grid(theta, (rotation.shape[0], 1, grid_size, grid_size, grid_size), align_corners=False)
        return grid.permute(0, 4, 1, 2, 3).contiguous()
    _require_inference(nn.Identity(), rotation, translation)
    theta = _affine_3x4(rotation, translation, invert)
    em0 = torch.zeros((rotation.shape[0], 1, 1, 1, 3), device=rotation.device, dtype=torch.float32)
    return ops.warp_field(em0, theta, grid_size)


class _WarpGenerator(nn.Module):
    _invert = False

    def __init__(self, num_channels):
        super().__init__()
        self.flowfield = FlowField()
        self.num_channels = COMPRESS_DIM
        # the reference's `nn.Parameter(...).to(device)` registers these only on CPU builds (model.py:934-935);
        # here they are always registered so they follow .to()/state_dict() on every device.
        self.adaptive_matrix_gamma = nn.Parameter(torch.randn(self.num_channels, self.num_channels))
        self.adaptive_matrix_beta = nn.Parameter(torch.randn(self.num_channels, self.num_channels))

    def _em_theta(self, R, t, z, e) -> Tuple[torch.Tensor, torch.Tensor]:
        assert R.shape == (z.shape[0], 3), f"Expected R shape (batch_size, 3), got {R.shape}"
        assert t.shape == (z.shape[0], 3), f"Expected t shape (batch_size, 3), got {t.shape}"
        assert z.shape == e.shape, f"Expected z and es to have the same shape, got {z.shape} and {e.shape}"
        s = torch.matmul((z + e).float(), self.adaptive_matrix_gamma.detach().float())
        em = self.flowfield._forward_cl(s)
        return em, _affine_3x4(R, t, self._invert)

    def _forward_autograd(self, R, t, z, e):
        """Differentiable form (row f-2): the flow-field tower through libmpb200; the 3-channel rigid grid (`F.affine_grid`) and the
        trilinear 16^3 -> 64^3 resize of the flow (model.py:965-973) through ATen's autograd."""
        assert R.shape == (z.shape[0], 3), f"Expected R shape (batch_size, 3), got {R.shape}"
        assert t.shape == (z.shape[0], 3), f"Expected t shape (batch_size, 3), got {t.shape}"
        assert z.shape == e.shape, f"Expected z and es to have the same shape, got {z.shape} and {e.shape}"
        s = torch.matmul((z + e).float(), self.adaptive_matrix_gamma.float())
        em = self.flowfield._forward_autograd(s[:, :, None, None])
        theta = _affine_3x4(R, t, self._invert)
        w_rt = F.affine_grid(theta, (z.shape[0], 1, 64, 64, 64), align_corners=False).permute(0, 4, 1, 2, 3)
        return w_rt + F.interpolate(em, size=(64, 64, 64), mode="trilinear", align_corners=False)

    def _forward(self, R, t, z, e):
        if _wants_grad(self, R, t, z, e):
            _require_inference(self, *(v.detach() for v in (R, t, z, e)))
            return _public(self._forward_autograd(R, t, z, e))
        _require_inference(self, R, t, z, e)
        em, theta = self._em_theta(R, t, z, e)
        return ops.warp_field(em, theta, 64)


class WarpGeneratorS2C(_WarpGenerator):
    """model.py:927-975."""
    _invert = True

    def forward(self, Rs, ts, zs, es):
        return self._forward(Rs, ts, zs, es)


class WarpGeneratorC2D(_WarpGenerator):
    """model.py:978-1024."""
    _invert = False

    def forward(self, Rd, td, zd, es):
        return self._forward(Rd, td, zd, es)


def apply_warping_field(v, warp_field):
    """model.py:1028-1065.  Differentiable (row f-2, first operator): with autograd recording on and an input that requires
    grad it goes through `ops.WarpFunction` (CUDA backward for both v and warp_field)."""
    for t in (v, warp_field):
        if torch.is_tensor(t) and not t.is_cuda:
            raise RuntimeError(f"apply_warping_field: input is on {t.device}; the B200 path has no CPU fallback")
    if torch.is_grad_enabled() and (v.requires_grad or warp_field.requires_grad):
        return ops.WarpFunction.apply(v, warp_field)
    return ops.apply_warping_field_ncdhw(_as_f32_cuda(v), _as_f32_cuda(warp_field))


# ----------------------------------------------------------------------------------------------------- pyramid
class AntiAliasInterpolation2d(nn.Module):
    """model.py:646-691."""

    def __init__(self, channels, scale):
        super().__init__()
        sigma = (1 / scale - 1) / 2
        kernel_size = 2 * round(sigma * 4) + 1
        self.ka = kernel_size // 2
        self.kb = self.ka - 1 if kernel_size % 2 == 0 else self.ka
        ax = torch.arange(kernel_size, dtype=torch.float32)
        mean = (kernel_size - 1) / 2
        g = torch.exp(-(ax - mean) ** 2 / (2 * sigma ** 2))
        kernel = g[:, None] * g[None, :]
        kernel = kernel / torch.sum(kernel)
        kernel = kernel.view(1, 1, kernel_size, kernel_size).repeat(channels, 1, 1, 1)
        self.register_buffer('weight', kernel)
        self.groups = channels
        self.scale = scale

    def forward(self, input):
        if self.scale == 1.0:
            return input
        if torch.is_grad_enabled() and input.requires_grad:
            # differentiable form (row f-2): a 3-channel depthwise blur + nearest sub-sampling (0.3 GFLOP) through ATen's autograd
            out = F.conv2d(F.pad(input, (self.ka, self.kb, self.ka, self.kb)), weight=self.weight, groups=self.groups)
            return F.interpolate(out, scale_factor=(self.scale, self.scale))
        _require_inference(self, input)
        step = int(round(1.0 / self.scale))
        return ops.blur_subsample(_as_f32_cuda(input), self.weight[0, 0].contiguous(), step)


class ImagePyramide(torch.nn.Module):
    """model.py:1070-1085."""

    def __init__(self, scales, num_channels):
        super().__init__()
        downs = {}
        for scale in scales:
            downs[str(scale).replace('.', '-')] = AntiAliasInterpolation2d(num_channels, scale)
        self.downs = nn.ModuleDict(downs)

    def forward(self, x):
        out_dict = {}
        for scale, down_module in self.downs.items():
            out_dict['prediction_' + str(scale).replace('-', '.')] = down_module(x)
        return out_dict


# ----------------------------------------------------------------------------------------------------- Gbase
class Gbase(nn.Module):
    """model.py:1127-1180.  `forward(xs, xd) -> (xhat_base, pyramids)`.

    Extra (non-reference) entry points for "1 source x N drivers" (BASELINE configs 2-3): `encode_source(xs)` runs the
    source-only half once, `drive(src, xd)` the per-driver half; `forward` is exactly `drive(encode_source(xs), xd)`.
    """

    def __init__(self):
        super().__init__()
        self.appearanceEncoder = Eapp()
        self.motionEncoder = Emtn()
        self.warp_generator_s2c = WarpGeneratorS2C(num_channels=512)
        self.warp_generator_c2d = WarpGeneratorC2D(num_channels=512)
        self.G3d = G3d(in_channels=96)
        self.G2d = G2d(in_channels=96)
        self.image_pyramid = ImagePyramide(scales=[0.5, 0.25], num_channels=3)
        self.tf32_motion = True   # only used when motionEncoder.backend == "cudnn" (stock incumbent path)

    def _emtn(self, x):
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=bool(self.tf32_motion)):
            return self.motionEncoder(x)

    def encode_source(self, xs, keep_stages: bool = False) -> Dict[str, object]:
        """Source-only half (model.py:1141-1160): Eapp, Emtn(xs), S2C warp, G3d.  Returns the cached state."""
        with torch.no_grad():       # non-reference, inference-only entry point: never records autograd
            return self._encode_source(xs, keep_stages)

    def _encode_source(self, xs, keep_stages: bool = False) -> Dict[str, object]:
        _require_inference(self, xs)
        if self.training:
            raise NotImplementedError("Gbase: train mode is not implemented on the B200 path (SURVEY.md 8f-2)")
        xs = _as_f32_cuda(xs)
        B = xs.shape[0]
        if _SOURCE_FORK and B <= 4 and xs.is_cuda and ops.PROFILE is None:
            # Small source batches are latency-bound (~400 launches of 8-64 CTAs each): the three branches that only read the
            # frame -- appearance volume, ResNet-50 descriptor, motion encoder -- run on three streams and meet before the S2C
            # warp generator.  (Under CUDA-graph capture the side streams become parallel branches of the graph.)
            cur = torch.cuda.current_stream(xs.device)
            s1, s2 = _side_streams(xs.device)
            s1.wait_stream(cur)
            s2.wait_stream(cur)
            with torch.cuda.stream(s1):
                es = self.appearanceEncoder._descriptor(xs)
                es.record_stream(cur)
            with torch.cuda.stream(s2):
                Rs, ts, zs = self._emtn(xs)
                for t in (Rs, ts, zs):
                    t.record_stream(cur)
            vs = self.appearanceEncoder._volume_cl(xs)
            cur.wait_stream(s1)
            cur.wait_stream(s2)
        else:
            with ops.stage("Eapp.volume", 866.5e9 * B):
                vs = self.appearanceEncoder._volume_cl(xs)
            with ops.stage("Eapp.descriptor", 34.4e9 * B):
                es = self.appearanceEncoder._descriptor(xs)
            with ops.stage("Emtn(source)", 235.6e9 * B):
                Rs, ts, zs = self._emtn(xs)
        with ops.stage("WarpGeneratorS2C", 0.5e9 * B):
            em, theta = self.warp_generator_s2c._em_theta(Rs, ts, zs, es)
        with ops.stage("warp(vs)+G3d", 163.1e9 * B):
            vc = ops.warp_fused(vs, em, theta, sum_d=False, f32=True, split=True)
            assert vc.shape[1:] == (16, 64, 64, 96), f"Expected vc shape (_, 96, 16, 64, 64), got {vc.shape}"
            vc2d = self.G3d._forward_cl(vc)
        src = {"vc2d": vc2d, "es": es}
        if keep_stages:
            src.update(vs=vs, Rs=Rs, ts=ts, zs=zs, em_s2c=em, theta_s2c=theta, vc=vc)
        return src

    def drive_motion(self, xd):
        """Driver-only part of the per-driver half (model.py:1145): Emtn(xd) -> (Rd, td, zd).  Independent of the
        source, so it can run concurrently with `encode_source` (engine.GraphedGbase does)."""
        with torch.no_grad():
            return self._drive_motion(xd)

    def _drive_motion(self, xd):
        _require_inference(self, xd)
        with ops.stage("Emtn(drivers)", 235.6e9 * xd.shape[0]):
            return self._emtn(_as_f32_cuda(xd))

    def drive_render(self, src: Dict[str, object], motion, keep_stages: bool = False):
        """C2D warp generator, fused warp + depth sum, G2d, pyramid (model.py:1163-1180)."""
        with torch.no_grad():
            return self._drive_render(src, motion, keep_stages)

    def _drive_render(self, src: Dict[str, object], motion, keep_stages: bool = False):
        Rd, td, zd = motion
        n = zd.shape[0]
        es = src["es"]
        if es.shape[0] != n:
            assert es.shape[0] == 1, f"source batch {es.shape[0]} does not match driver batch {n}"
            es = es.expand(n, -1)
        with ops.stage("WarpGeneratorC2D", 0.5e9 * n):
            em, theta = self.warp_generator_c2d._em_theta(Rd, td, zd, es)
        with ops.stage("warp(vc2d)+sumD"):
            proj = ops.warp_fused(src["vc2d"], em, theta, sum_d=True, f32=keep_stages, split=True)
        assert proj.shape[1:] == (1, 64, 64, 96), f"Expected vc2d_warped shape (_, 96, 16, 64, 64), got {proj.shape}"
        with ops.stage("G2d", 504.6e9 * n):
            xhat = self.G2d._forward_cl(proj)
        with ops.stage("ImagePyramide", 0.31e9 * n):
            pyramids = self.image_pyramid(xhat)
        if keep_stages:
            return xhat, pyramids, dict(Rd=Rd, td=td, zd=zd, em_c2d=em, theta_c2d=theta, projected=proj)
        return xhat, pyramids

    def drive(self, src: Dict[str, object], xd, keep_stages: bool = False):
        """Per-driver half (model.py:1145, 1163-1180).  `src["vc2d"]` may hold 1 sample (shared source) or len(xd)."""
        return self.drive_render(src, self.drive_motion(xd), keep_stages)

    def forward(self, xs, xd):
        """model.py:1140-1180.  Inference-only: with autograd recording on and trainable parameters it raises
        NotImplementedError rather than returning detached outputs.

        The ~470 kernels of a forward cost about as much host time to enqueue as they take on the GPU at small batch, so the
        drop-in entry point replays a CUDA graph: the second call with a given (batch size, device) captures one, later calls
        copy the inputs into its static buffers, replay, and return CLONES of the outputs (callers may keep them).  The
        graph is keyed on the weights' signature (`_sig`): any parameter update, `.to()`, `load_state_dict` or
        `invalidate_plans` drops it.  `self.forward_graphs = False` (or MPB200_FORWARD_GRAPHS=0) launches eagerly."""
        assert xs.shape[0] == xd.shape[0], f"Expected zs and es to have the same shape (Bs == Bd), got {xs.shape[0]} and {xd.shape[0]}"
        if self.training or _wants_grad(self, xs, xd):
            _require_inference(self, xs.detach(), xd.detach())
            return _public(self._forward_autograd(xs.float(), xd.float()))
        if (_FWD_GRAPHS and getattr(self, "forward_graphs", True) and not self.training and xs.is_cuda and xd.is_cuda
                and xs.device == xd.device and tuple(xs.shape[1:]) == (3, 512, 512) and tuple(xd.shape[1:]) == (3, 512, 512)
                and not _capturing(xs)):
            return self._forward_graphed(xs, xd)
        return self._forward_eager(xs, xd)

    def _forward_autograd(self, xs, xd):
        """Differentiable / train-mode `forward` (row f-2; train.py:194, 283): the reference's statement order (model.py:1141-1178)
        on the autograd Functions of ops.py -- every convolution (forward, data and weight gradient), GroupNorm, train-mode
        BatchNorm and both `apply_warping_field` calls run on libmpb200; pools, resizes and the rigid 3-channel grid go
        through ATen's autograd."""
        assert tuple(xs.shape[1:]) == (3, 512, 512) and tuple(xd.shape[1:]) == (3, 512, 512), \
            f"Expected input shape (_, 3, 512, 512), got {tuple(xs.shape)}"
        vs, es = self.appearanceEncoder._forward_autograd(xs)
        Rs, ts, zs = self.motionEncoder._forward_autograd(xs)
        Rd, td, zd = self.motionEncoder._forward_autograd(xd)
        w_s2c = self.warp_generator_s2c._forward_autograd(Rs, ts, zs, es)
        vc = ops.WarpFunction.apply(vs, w_s2c)
        assert vc.shape[1:] == (96, 16, 64, 64), f"Expected vc shape (_, 96, 16, 64, 64), got {vc.shape}"
        vc2d = self.G3d._forward_autograd(vc)
        w_c2d = self.warp_generator_c2d._forward_autograd(Rd, td, zd, es)
        vc2d_warped = ops.WarpFunction.apply(vc2d, w_c2d)
        assert vc2d_warped.shape[1:] == (96, 16, 64, 64), f"Expected vc2d_warped shape (_, 96, 16, 64, 64), got {vc2d_warped.shape}"
        xhat_base = self.G2d._forward_autograd(torch.sum(vc2d_warped, dim=2))
        return xhat_base, self.image_pyramid(xhat_base)

    def _forward_eager(self, xs, xd):
        src = self.encode_source(xs)
        return self.drive(src, xd)

    def release_forward_graphs(self) -> None:
        """Drop the CUDA graphs `forward` has captured (each owns the activation memory of its batch size)."""
        self.__dict__.pop("_mp_fwd_graphs", None)
        self.__dict__.pop("_mp_fwd_seen", None)

    def _graph_sig(self):
        rot = self.motionEncoder.rotation_net.model
        return (_sig(self), tuple((t.data_ptr(), t._version) for t in list(rot.parameters()) + list(rot.buffers())),
                self.motionEncoder.backend, self.appearanceEncoder.custom_resnet50.backend)

    def _forward_graphed(self, xs, xd):
        key = (int(xs.shape[0]), str(xs.device))
        sig = self._graph_sig()
        cache = self.__dict__.setdefault("_mp_fwd_graphs", {})
        ent = cache.get(key)
        if ent is not None and ent["sig"] != sig:
            cache.pop(key)
            ent = None
        if ent is None:
            seen = self.__dict__.setdefault("_mp_fwd_seen", {})
            seen[key] = seen.get(key, 0) + 1
            if seen[key] < 2:                      # first sighting: eager (it also packs the plans and calibrates the FP8 scales)
                return self._forward_eager(xs, xd)
            while len(cache) >= 2:                 # a graph owns the activations of its batch size: keep at most two
                cache.pop(next(iter(cache)))
            dev = xs.device
            xs_b = torch.empty_like(_as_f32_cuda(xs))
            xd_b = torch.empty_like(_as_f32_cuda(xd))
            xs_b.copy_(xs)
            xd_b.copy_(xd)
            cur = torch.cuda.current_stream(dev)
            side = torch.cuda.Stream(dev)
            side.wait_stream(cur)
            with torch.cuda.stream(side), torch.no_grad():
                self._forward_eager(xs_b, xd_b)     # allocator / plans warm on the capture stream's pool
            cur.wait_stream(side)
            torch.cuda.synchronize(dev)
            g = torch.cuda.CUDAGraph()
            l0 = ops.LAUNCHES
            with torch.cuda.graph(g), torch.no_grad():
                out = self._forward_eager(xs_b, xd_b)
            ent = {"sig": sig, "graph": g, "xs": xs_b, "xd": xd_b, "out": out, "launches": ops.LAUNCHES - l0}
            cache[key] = ent
            seen.pop(key, None)
        ent["xs"].copy_(xs, non_blocking=True)
        ent["xd"].copy_(xd, non_blocking=True)
        ent["graph"].replay()
        ops._count(ent["launches"])
        img, pyr = ent["out"]
        return img.clone(), {k: v.clone() for k, v in pyr.items()}


ADAPTIVE_KEYS = tuple(f"warp_generator_{g}.adaptive_matrix_{m}" for g in ("s2c", "c2d") for m in ("gamma", "beta"))


def load_reference_state_dict(gbase: "Gbase", state_dict, adaptive_matrices=None):
    """`Gbase.load_state_dict` for checkpoints written by the reference.  On CUDA builds the reference's
    `nn.Parameter(...).to(device)` (model.py:934-935, 985-986) leaves `adaptive_matrix_gamma/beta` OUT of `state_dict()`, so
    such a checkpoint lacks exactly these 4 keys (`ADAPTIVE_KEYS`); every other key must be present.  Pass the matrices in
    `adaptive_matrices` ({key: tensor}) when the training run kept them; otherwise they stay at this module's init and a
    warning says so (the output then depends on this process's random state -- seed it or supply the matrices)."""
    import warnings
    sd = dict(state_dict)
    if adaptive_matrices:
        sd.update({k: v for k, v in adaptive_matrices.items() if k in ADAPTIVE_KEYS})
    res = gbase.load_state_dict(sd, strict=False)
    missing = [k for k in res.missing_keys if not k.startswith("image_pyramid.")]
    bad = [k for k in missing if k not in ADAPTIVE_KEYS]
    if bad or res.unexpected_keys:
        raise RuntimeError(f"load_reference_state_dict: missing {bad[:5]}{'...' if len(bad) > 5 else ''}, "
                           f"unexpected {list(res.unexpected_keys)[:5]}")
    if missing:
        warnings.warn("checkpoint has no adaptive_matrix_gamma/beta (written by a CUDA build of the reference, model.py:934-935): "
                      f"{missing} keep this module's random init -- pass `adaptive_matrices=` or seed the construction")
    invalidate_plans(gbase)
    return res


# ----------------------------------------------------------------------------------------------------- Genh / GHR
class Genh(nn.Module, _Packed):
    """model.py:1346-1391: 7x7 stem, 4 ResBlock2D with AvgPool2d between them, 8 ResBlock2D, 3 x
    [bilinear x2, ResBlock2D], 7x7 conv to RGB, tanh -- fully convolutional, output size = input size.

    One documented repair: the reference writes `ResBlock2D(64)` although `ResBlock2D.__init__(in_channels,
    out_channels, downsample=False)` (model.py:601) has no default, so `Genh()` raises TypeError there (model.py:1354);
    the only reading that type-checks and keeps the residual identity is `ResBlock2D(64, 64)`.  Attribute names and
    state_dict keys are the reference's (`encoder.0.weight`, `encoder.1.conv1.weight`, ..., `decoder.6.weight`)."""

    def __init__(self):
        super().__init__()
        rb = lambda: ResBlock2D(64, 64)
        self.encoder = nn.Sequential(nn.Conv2d(3, 64, kernel_size=7, padding=3), rb(), nn.AvgPool2d(kernel_size=2, stride=2),
                                     rb(), nn.AvgPool2d(kernel_size=2, stride=2), rb(),
                                     nn.AvgPool2d(kernel_size=2, stride=2), rb())
        self.res_blocks = nn.Sequential(*[rb() for _ in range(8)])
        up = lambda: nn.Upsample(scale_factor=2, mode='bilinear', align_corners=True)
        self.decoder = nn.Sequential(up(), rb(), up(), rb(), up(), rb(), nn.Conv2d(64, 3, kernel_size=7, padding=3), nn.Tanh())

    def _build_plan(self):
        dev = self.encoder[0].weight.device
        return {"stem": ops.pack_conv(self.encoder[0].weight, self.encoder[0].bias, dev, cin_pad=16),
                "out": ops.pack_conv(self.decoder[6].weight, self.decoder[6].bias, dev)}

    def _forward_cl(self, x: torch.Tensor) -> torch.Tensor:
        """x NCHW fp32 [B,3,H,W] (H, W multiples of 8) -> NCHW fp32 [B,3,H,W] in (-1, 1)."""
        from .emtn_cuda import rgb16
        P = self._plan()
        h, _ = ops.conv(rgb16(x), P["stem"], f32=False, split=True)
        for i in (1, 3, 5, 7):
            pooled = i < 7
            h, _ = self.encoder[i]._forward_cl(h, f32=pooled, split=not pooled)
            if pooled:
                h = ops.avgpool2(h, 1, f32=False, split=True)
        for blk in self.res_blocks:
            h, _ = blk._forward_cl(h)
        for i in (1, 3, 5):
            h = ops.upsample2x_linear(h, 1, f32=False, split=True)
            h, _ = self.decoder[i]._forward_cl(h)
        out, _ = ops.conv(h, P["out"], act=ACT_TANH, f32=True)
        return ops.to_nchw(out, 4)

    def forward(self, x):
        _require_inference(self, x)
        if self.training:
            raise NotImplementedError("Genh: train mode is not implemented on the B200 path (SURVEY.md 8f-2)")
        assert x.dim() == 4 and x.shape[1] == 3 and x.shape[2] % 8 == 0 and x.shape[3] % 8 == 0, \
            f"Genh expects (B, 3, H, W) with H, W multiples of 8 (three 2x poolings), got {tuple(x.shape)}"
        with torch.no_grad():
            return self._forward_cl(_as_f32_cuda(x))


class GHR(nn.Module):
    """model.py:1441-1450: `Genh(Gbase(xs, xd))`.  The reference hands Genh the whole `(image, pyramids)` tuple that
    `Gbase.forward` returns (model.py:1180 vs 1447-1448) -- a TypeError as written; here the
    image is taken, which is the only reading under which `README.md:214` (`GHR.Gbase.load_state_dict(...)`) and
    `GHR.forward` run."""

    def __init__(self):
        super().__init__()
        self.Gbase = Gbase()
        self.Genh = Genh()

    def forward(self, xs, xd):
        xhat_base = self.Gbase(xs, xd)
        xhat_base = xhat_base[0] if isinstance(xhat_base, tuple) else xhat_base
        return self.Genh(xhat_base)
