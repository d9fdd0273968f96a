"""Tensor-level wrappers over the C ABI (include/mpb200.h).  PyTorch is used for device memory and streams only.

Internal activation format (`Act`): channels-last [N, D, H, W, C] (2-D tensors carry D = 1), held as an fp32 tensor
and/or a pair of bf16 planes (hi, lo) with x ~= hi + lo -- the operand format of the split-bf16 tensor-core
convolution.  The pooled motion-encoder trunks (Emtn) use a third form, one fp16 plane `h16`, consumed by the
two-pass `PREC_F16X2` convolution.  Every function launches on the current torch CUDA stream and never synchronises.
"""
from __future__ import annotations

import ctypes
import math
import os
from typing import Optional, Sequence, Tuple

import torch

from . import lib as _lib
from .lib import (ACT_NONE, ACT_RELU, ACT_RELU_TANH, ACT_SIGMOID, ACT_TANH, FMT_F16_Q8, FMT_SPLIT_BF16, PREC_F16_Q8,  # noqa: F401
                  PREC_F16X2, PREC_SPLIT_BF16, ConvDesc)

LAUNCHES = 0   # number of libmpb200 kernel launches issued by this process (bench.py reports it)
PROFILE = None  # when a list: (kind, start_event, end_event, algorithmic_flops, algorithmic_bytes) per hot launch


class _Prof:
    """CUDA-event bracket on the launching stream, active only while `PROFILE` is a list (bench.py roofline leg)."""
    __slots__ = ("kind", "flops", "bytes", "e0")

    def __init__(self, kind, flops=0, nbytes=0):
        self.kind, self.flops, self.bytes, self.e0 = kind, flops, nbytes, None

    def __enter__(self):
        if PROFILE is not None:
            self.e0 = torch.cuda.Event(enable_timing=True)
            self.e0.record()
        return self

    def __exit__(self, *exc):
        if self.e0 is not None:
            e1 = torch.cuda.Event(enable_timing=True)
            e1.record()
            PROFILE.append((self.kind, self.e0, e1, self.flops, self.bytes))
        return False


class stage:
    """`with ops.stage("G3d"):` -- stage-level CUDA-event bracket, recorded only while `PROFILE` is a list."""

    def __init__(self, name: str, flops: float = 0.0):
        self.p = _Prof("stage:" + name, flops)

    def __enter__(self):
        self.p.__enter__()
        return self

    def __exit__(self, *exc):
        return self.p.__exit__(*exc)


def _profiled(fn):
    """Bracket a wrapper with CUDA events while `PROFILE` is a list (per-launch dump of bench.py)."""
    import functools

    @functools.wraps(fn)
    def wrapper(*a, **k):
        if PROFILE is None:
            return fn(*a, **k)
        with _Prof(fn.__name__):
            return fn(*a, **k)
    return wrapper


def _count(n: int = 1) -> None:
    global LAUNCHES
    LAUNCHES += n


def _stream() -> int:
    return torch.cuda.current_stream().cuda_stream


def _p(t: Optional[torch.Tensor]):
    return None if t is None else t.data_ptr()


def _chk_cuda(t: torch.Tensor, dtype, what: str) -> None:
    if not t.is_cuda:
        raise RuntimeError(f"{what}: libmpb200 kernels need a CUDA tensor (got {t.device}); there is no CPU fallback")
    if t.dtype != dtype:
        raise RuntimeError(f"{what}: expected {dtype}, got {t.dtype}")
    if not t.is_contiguous():
        raise RuntimeError(f"{what}: tensor must be contiguous")


class Act:
    """Channels-last activation [N, D, H, W, C]: `f32` and/or the split pair (`hi`, `lo`) and/or one fp16 plane `h16`
    (optionally with the FP8 byte plane `q8` [N, D, H, W, 2C] of the "F16_Q8" operand format, include/mpb200.h, written
    with the per-tensor power-of-two scale `q8_scale`)."""
    __slots__ = ("f32", "hi", "lo", "h16", "q8", "shape", "q8_scale")

    def __init__(self, shape, f32=None, hi=None, lo=None, h16=None, q8=None, q8_scale=1.0):
        self.shape = tuple(int(s) for s in shape)
        self.f32, self.hi, self.lo, self.h16, self.q8 = f32, hi, lo, h16, q8
        self.q8_scale = float(q8_scale)

    @property
    def N(self): return self.shape[0]
    @property
    def D(self): return self.shape[1]
    @property
    def H(self): return self.shape[2]
    @property
    def W(self): return self.shape[3]
    @property
    def C(self): return self.shape[4]
    @property
    def S(self): return self.shape[1] * self.shape[2] * self.shape[3]

    @property
    def device(self):
        return next(t for t in (self.f32, self.hi, self.h16) if t is not None).device

    def has_split(self):
        return self.hi is not None


def _alloc(shape, device, f32: bool, split: bool, h16: bool = False, q8: bool = False):
    a = Act(shape)
    if q8:
        a.q8 = torch.empty(tuple(shape[:-1]) + (2 * shape[-1],), dtype=torch.uint8, device=device)
    if f32:
        a.f32 = torch.empty(shape, dtype=torch.float32, device=device)
    if split:
        a.hi = torch.empty(shape, dtype=torch.bfloat16, device=device)
        a.lo = torch.empty(shape, dtype=torch.bfloat16, device=device)
    if h16:
        a.h16 = torch.empty(shape, dtype=torch.float16, device=device)
    return a


# ----------------------------------------------------------------------------------------------------- layout
@_profiled
def from_nchw(x: torch.Tensor, f32: bool = False, split: bool = True) -> Act:
    """NCHW / NCDHW fp32 -> Act."""
    _chk_cuda(x, torch.float32, "from_nchw")
    if x.dim() == 4:
        N, C, H, W = x.shape
        D = 1
    else:
        N, C, D, H, W = x.shape
    out = _alloc((N, D, H, W, C), x.device, f32, split)
    L = _lib.load()
    _lib.check(L.mp_nchw_to_cl(_p(x), _p(out.f32), _p(out.hi), _p(out.lo), N, C, D * H * W, _stream()), "mp_nchw_to_cl")
    _count()
    return out


@_profiled
def from_nchw_pad16(x: torch.Tensor) -> Act:
    """NCHW fp32 [N,C<=16,H,W] -> split channels-last activation with 16 (zero-padded) channels."""
    _chk_cuda(x, torch.float32, "from_nchw_pad16")
    N, C, H, W = x.shape
    out = _alloc((N, 1, H, W, 16), x.device, False, True)
    L = _lib.load()
    _lib.check(L.mp_nchw_to_cl_pad16(_p(x), _p(out.hi), _p(out.lo), N, C, H * W, _stream()), "mp_nchw_to_cl_pad16")
    _count()
    return out


@_profiled
def to_nchw(a: Act, ndim: int = 5) -> torch.Tensor:
    N, D, H, W, C = a.shape
    out = torch.empty((N, C, D, H, W) if ndim == 5 else (N, C, H, W), dtype=torch.float32, device=a.device)
    L = _lib.load()
    _lib.check(L.mp_cl_to_nchw(_p(a.f32), _p(a.hi), _p(a.lo), _p(out), N, C, D * H * W, _stream()), "mp_cl_to_nchw")
    _count()
    return out


@_profiled
def ensure_split(a: Act) -> Act:
    if a.hi is None:
        a.hi = torch.empty(a.shape, dtype=torch.bfloat16, device=a.device)
        a.lo = torch.empty(a.shape, dtype=torch.bfloat16, device=a.device)
        L = _lib.load()
        _lib.check(L.mp_split(_p(a.f32), _p(a.hi), _p(a.lo), a.f32.numel(), _stream()), "mp_split")
        _count()
    return a


@_profiled
def avgpool2(a: Act, pool_d: int, f32: bool = True, split: bool = False) -> Act:
    N, D, H, W, C = a.shape
    out = _alloc((N, D // pool_d, H // 2, W // 2, C), a.device, f32, split)
    L = _lib.load()
    _lib.check(L.mp_avgpool2_cl(_p(a.f32), _p(out.f32), _p(out.hi), _p(out.lo), N, D, H, W, C, pool_d, _stream()),
               "mp_avgpool2_cl")
    _count()
    return out


@_profiled
def upsample2x_linear(a: Act, up_d: int, f32: bool = False, split: bool = True) -> Act:
    N, D, H, W, C = a.shape
    out = _alloc((N, D * up_d, H * 2, W * 2, C), a.device, f32, split)
    L = _lib.load()
    _lib.check(L.mp_upsample2x_linear_cl(_p(a.f32), _p(a.hi), _p(a.lo), _p(out.f32), _p(out.hi), _p(out.lo), N, D, H, W,
                                         C, up_d, _stream()), "mp_upsample2x_linear_cl")
    _count()
    return out


@_profiled
def upsample2x_bilinear_hq(a: Act, q8_scale: float = 1.0) -> Act:
    """nn.Upsample(x2, bilinear, align_corners=True) from split planes to the F16_Q8 plane pair (2-D, C % 64 == 0)."""
    N, D, H, W, C = a.shape
    if D != 1 or a.hi is None:
        raise RuntimeError("upsample2x_bilinear_hq: needs a 2-D split-bf16 activation")
    out = _alloc((N, 1, H * 2, W * 2, C), a.device, False, False, True, True)
    out.q8_scale = float(q8_scale)
    L = _lib.load()
    _lib.check(L.mp_upsample2x_bilinear_hq(_p(a.hi), _p(a.lo), _p(out.h16), _p(out.q8), N, H, W, C, float(q8_scale), _stream()),
               "mp_upsample2x_bilinear_hq")
    _count()
    return out


@_profiled
def upsample_nearest(a: Act, scale: Sequence[int], f32: bool = False, split: bool = True) -> Act:
    N, D, H, W, C = a.shape
    sd, sh, sw = scale
    out = _alloc((N, D * sd, H * sh, W * sw, C), a.device, f32, split)
    L = _lib.load()
    _lib.check(L.mp_upsample_nearest_cl(_p(a.f32), _p(out.f32), _p(out.hi), _p(out.lo), N, D, H, W, C, sd, sh, sw,
                                        _stream()), "mp_upsample_nearest_cl")
    _count()
    return out


# ----------------------------------------------------------------------------------------------------- norm
def new_stats(N: int, G: int, device) -> torch.Tensor:
    return torch.zeros((N, G, 2), dtype=torch.float64, device=device)


@_profiled
def gn_stats(a: Act, G: int) -> torch.Tensor:
    st = new_stats(a.N, G, a.device)
    L = _lib.load()
    _lib.check(L.mp_gn_stats(_p(a.f32), _p(st), a.N, a.S, a.C, G, _stream()), "mp_gn_stats")
    _count()
    return st


@_profiled
def gn_finalize(stats: torch.Tensor, a_shape, G: int, gamma=None, beta=None, gamma2=None, beta2=None,
                eps: float = 1e-5) -> torch.Tensor:
    N, D, H, W, C = a_shape
    ab = torch.empty((N, C, 2), dtype=torch.float32, device=stats.device)
    L = _lib.load()
    _lib.check(L.mp_gn_finalize(_p(stats), _p(gamma), _p(beta), _p(gamma2), _p(beta2), _p(ab), N, D * H * W, C, G,
                                eps, _stream()), "mp_gn_finalize")
    _count()
    return ab


@_profiled
def affine_act(a: Act, ab: Optional[torch.Tensor], res: Optional[Act] = None, act: int = ACT_NONE, f32: bool = False,
               split: bool = True) -> Act:
    out = _alloc(a.shape, a.device, f32, split)
    rf = res.f32 if res is not None else None
    rh = res.hi if (res is not None and rf is None) else None
    rl = res.lo if (res is not None and rf is None) else None
    L = _lib.load()
    _lib.check(L.mp_affine_act_cl(_p(a.f32), _p(ab), _p(rf), _p(rh), _p(rl), _p(out.f32), _p(out.hi), _p(out.lo), a.N,
                                  a.S, a.C, act, _stream()), "mp_affine_act_cl")
    _count()
    return out


def group_norm_act(a: Act, G: int, stats: Optional[torch.Tensor] = None, gamma=None, beta=None, gamma2=None,
                   beta2=None, res: Optional[Act] = None, act: int = ACT_RELU, f32: bool = False,
                   split: bool = True) -> Act:
    """act(GroupNorm_G(a) [*gamma2 + beta2] [+ res]); `stats` come from a conv epilogue when available."""
    if stats is None:
        stats = gn_stats(a, G)
    ab = gn_finalize(stats, a.shape, G, gamma, beta, gamma2, beta2)
    return affine_act(a, ab, res, act, f32, split)


# ----------------------------------------------------------------------------------------------------- conv
class PackedConv:
    """Weights of one convolution in kernel format: split-bf16 [Cout_pad, taps*Cin] (tap-major, cin-minor) + bias."""
    __slots__ = ("w_hi", "w_lo", "bias", "Cin", "Cout", "Cout_pad", "k", "prec", "Cin2", "corr_scale", "acc_scale")

    def __init__(self, w_hi, w_lo, bias, Cin, Cout, Cout_pad, k, prec=PREC_SPLIT_BF16, Cin2=0, corr_scale=0.0,
                 acc_scale=1.0):
        self.w_hi, self.w_lo, self.bias = w_hi, w_lo, bias
        self.corr_scale = corr_scale   # PREC_F16_Q8: weight of the FP8 cross-term accumulator, 1 / 2048 (per unit input scale)
        self.acc_scale = acc_scale     # fp16 planes hold w / acc_scale (power of two): the kernel rescales the accumulator
        self.Cin, self.Cout, self.Cout_pad, self.k, self.prec = Cin, Cout, Cout_pad, tuple(k), prec
        self.Cin2 = Cin2          # > 0: rows are [taps*Cin | Cin2]: a 1x1 shortcut over a second source is fused in


def standardize_weight(w: torch.Tensor) -> torch.Tensor:
    """Conv2d_WS / Conv3D_WS weight transform (model.py:61-69, 77-86), evaluated in float64."""
    w = w.double()
    dims = tuple(range(1, w.dim()))
    w = w - w.mean(dim=dims, keepdim=True)
    std = w.reshape(w.size(0), -1).std(dim=1).view(-1, *([1] * (w.dim() - 1))) + 1e-5
    return w / std


def fold_bn(w: torch.Tensor, b: Optional[torch.Tensor], bn: dict, eps: float = 1e-5):
    """conv -> BatchNorm(eval) == conv with w*s, (b-mean)*s+beta, s = gamma/sqrt(var+eps); float64 (model.py:605-616)."""
    s = bn["weight"].double() / torch.sqrt(bn["running_var"].double() + eps)
    w = w.double() * s.view(-1, *([1] * (w.dim() - 1)))
    b0 = b.double() if b is not None else torch.zeros_like(s)
    return w, (b0 - bn["running_mean"].double()) * s + bn["bias"].double()


F16_LO_SCALE = 2048.0   # PREC_F16X2: w_lo holds (w - fp16(w)) * 2^11 so that it stays in fp16's normal range


def e4m3(x: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-even, saturating float -> float8_e4m3fn (what `__nv_cvt_float2_to_fp8x2(.., SATFINITE, E4M3)` does)."""
    return x.float().clamp(-448.0, 448.0).to(torch.float8_e4m3fn)


def q8_scale_for(amax: float, target: float = 96.0) -> float:
    """Power-of-two scale that puts a tensor's largest magnitude near `target` (e4m3 keeps 3 mantissa bits from 2^-6 up
    to 448): ~4.6x head-room above the calibrated maximum, 12 binades below it."""
    if not (amax > 0.0) or not math.isfinite(amax):
        return 1.0
    return 2.0 ** round(math.log2(target / amax))


def q8_planes(x: torch.Tensor, scale: float = 1.0) -> torch.Tensor:
    """fp32 channels-last [..., C] (C % 64 == 0) -> the byte plane [..., 2C] of the F16_Q8 operand format, written with
    the per-tensor power-of-two `scale` (what the convolution epilogue does with `out_q8_scale`)."""
    h = x.clamp(-65504.0, 65504.0).to(torch.float16)
    a8 = e4m3(x * scale).view(torch.uint8)
    al8 = e4m3((x - h.float()) * (F16_LO_SCALE * scale)).view(torch.uint8)
    lead = x.shape[:-1]
    return torch.stack((a8.reshape(*lead, -1, 64), al8.reshape(*lead, -1, 64)), -2).reshape(*lead, -1)


def pack_conv(weight: torch.Tensor, bias: Optional[torch.Tensor], device=None, cin_pad: int = 0,
              prec: int = PREC_SPLIT_BF16, shortcut: Optional[Tuple[torch.Tensor, Optional[torch.Tensor]]] = None) -> PackedConv:
    """weight (Cout, Cin, [kd,] kh, kw) float32/64 -> PackedConv on `device`.  `cin_pad` zero-pads the input-channel
    axis (RGB stems run on the tensor-core kernel with their 3 channels padded to 16).  `prec` selects the operand
    format: split-bf16 planes, or fp16 hi + scaled fp16 lo for the two-pass convolution."""
    device = device or weight.device
    w = weight.detach().to(device=device, dtype=torch.float64)
    if w.dim() == 4:
        w = w.unsqueeze(2)
    if cin_pad and cin_pad > w.shape[1]:
        w = torch.cat([w, torch.zeros(w.shape[0], cin_pad - w.shape[1], *w.shape[2:], dtype=w.dtype, device=device)], 1)
    Cout, Cin, kd, kh, kw = w.shape
    wk = w.permute(0, 2, 3, 4, 1).reshape(Cout, kd * kh * kw * Cin)
    Cin2 = 0
    if shortcut is not None:
        # `shortcut` = (weight (Cout, Cin2, 1, 1[, 1]), bias): a 1x1 convolution of a second source added inside the
        # accumulator (ResBlock2D / down-sampling ResNet blocks): its columns extend K, its bias joins the bias
        ws, bs = shortcut
        ws = ws.detach().to(device=device, dtype=torch.float64).reshape(Cout, -1)
        Cin2 = ws.shape[1]
        wk = torch.cat([wk, ws], 1)
        if bs is not None:
            bias = bs if bias is None else bias.detach().to(device=device, dtype=torch.float64) + \
                bs.detach().to(device=device, dtype=torch.float64)
    wk = wk.to(torch.float32)
    Cout_pad = (Cout + 15) // 16 * 16
    if Cout_pad != Cout:
        wk = torch.cat([wk, torch.zeros(Cout_pad - Cout, wk.shape[1], device=device)], 0)
    corr_scale, acc_scale = 0.0, 1.0
    if prec in (PREC_F16_Q8, PREC_F16X2):
        # fp16 planes are formed from w * sw, sw the power of two that puts the largest weight into [1, 2): weights of
        # any magnitude keep fp16's full 11 bits (no subnormals); the kernel multiplies the accumulator by 1 / sw
        amax = float(wk.abs().max())
        sw = 2.0 ** (-math.floor(math.log2(amax))) if (amax > 0 and math.isfinite(amax)) else 1.0
        wk = wk * sw
        acc_scale = 1.0 / sw
    if prec == PREC_F16_Q8:
        # fp16 main plane + byte plane: per 64-wide K chunk [e4m3(wl * 2048) x 64 | e4m3(w) x 64]
        if wk.shape[1] % 64:
            raise RuntimeError("pack_conv: PREC_F16_Q8 needs input channels in multiples of 64")
        hi = wk.to(torch.float16)
        wl8 = e4m3((wk - hi.float()) * F16_LO_SCALE).view(torch.uint8)
        w8 = e4m3(wk).view(torch.uint8)
        q = torch.stack((wl8.view(Cout_pad, -1, 64), w8.view(Cout_pad, -1, 64)), 2).reshape(Cout_pad, -1)
        corr_scale = 1.0 / F16_LO_SCALE
        hi, lo = hi.contiguous().view(torch.uint8), q.contiguous()            # both [Cout_pad, 2K] bytes
    elif prec == PREC_F16X2:
        hi = wk.to(torch.float16)
        lo = ((wk - hi.float()) * F16_LO_SCALE).to(torch.float16)
    else:
        hi = wk.to(torch.bfloat16)
        lo = (wk - hi.float()).to(torch.bfloat16)
    b = None if bias is None else bias.detach().to(device=device, dtype=torch.float32).contiguous()
    planes = torch.stack((hi, lo)).contiguous()      # one allocation: the kernel then fetches [hi | lo] with ONE TMA load
    return PackedConv(planes[0], planes[1], b, Cin, Cout, Cout_pad, (kd, kh, kw), prec, Cin2, corr_scale, acc_scale)


def pack_stem3x3_f16(weight: torch.Tensor, bias: Optional[torch.Tensor], device=None) -> PackedConv:
    """3x3 RGB stem weight (Cout, 3, 3, 3) -> 1x1 PREC_F16X2 PackedConv over the 32 patch channels written by
    `im2col3x3_f16` (channel (kh*3+kw)*3 + c, zero-padded from 27 to 32)."""
    Cout, Cin, kh, kw = weight.shape
    assert (Cin, kh, kw) == (3, 3, 3), weight.shape
    w = weight.detach().double().permute(0, 2, 3, 1).reshape(Cout, 27)
    w = torch.cat([w, torch.zeros(Cout, 5, dtype=w.dtype, device=w.device)], 1).reshape(Cout, 32, 1, 1)
    return pack_conv(w, bias, device or weight.device, prec=PREC_F16X2)


def _conv_q8(a: Act, pw: PackedConv, res: Optional[Act], act: int, f32: bool, split: bool, stats_groups: int,
             hq: bool, out_q8_scale: float = 1.0, src2: Optional[Act] = None) -> Tuple[Act, Optional[torch.Tensor]]:
    """Convolutions around the F16_Q8 operand format (fp16 plane + FP8 byte plane): `pw.prec == PREC_F16_Q8` consumes it
    (fp16 main product + FP8 cross terms), `hq=True` produces it -- from either kind of convolution, so the format change
    rides on an epilogue.  Stride 1, no channel windows; the residual may be fp32, split or F16_Q8."""
    q8_in = pw.prec == PREC_F16_Q8
    if q8_in and (a.h16 is None or a.q8 is None):
        raise RuntimeError("conv: a PREC_F16_Q8 convolution needs an activation with fp16 + FP8 planes")
    if not q8_in and a.hi is None:
        ensure_split(a)
    if pw.Cin != a.C or (hq and split) or (pw.Cin2 and not q8_in):
        raise RuntimeError("conv: unsupported F16_Q8 configuration")
    if pw.Cin2:
        # fused 1x1 shortcut over a second F16_Q8 source on the same grid; both byte planes share one cross-term scale, so the
        # two sources must have been written with the same per-tensor scale
        if src2 is None or src2.q8 is None or src2.shape[:4] != a.shape[:4] or src2.C != pw.Cin2:
            raise RuntimeError("conv: the fused F16_Q8 shortcut needs a second F16_Q8 source on the same grid")
        if src2.q8_scale != a.q8_scale:
            raise RuntimeError("conv: the two sources of a fused F16_Q8 shortcut must share the byte-plane scale")
    N, D, H, W, C = a.shape
    out = _alloc((N, D, H, W, pw.Cout), a.device, f32, split, hq, hq)
    out.q8_scale = float(out_q8_scale) if hq else 1.0
    stats = new_stats(N, stats_groups, a.device) if stats_groups else None
    d = ConvDesc()
    d.w_hi, d.w_lo, d.bias, d.prec = _p(pw.w_hi), _p(pw.w_lo), _p(pw.bias), pw.prec
    d.corr_scale = pw.corr_scale / a.q8_scale if q8_in else 0.0      # the input plane's scale leaves with the weight's
    d.acc_scale = pw.acc_scale
    d.out_q8_scale = out.q8_scale
    d.in_hi, d.in_lo = (_p(a.h16), _p(a.q8)) if q8_in else (_p(a.hi), _p(a.lo))
    if res is not None:
        if res.shape != out.shape:
            raise RuntimeError(f"conv: residual shape {res.shape} != output shape {out.shape}")
        if res.f32 is not None:
            d.res_f32 = _p(res.f32)
        elif res.q8 is not None:
            d.res_hi, d.res_lo, d.res_fmt, d.res_q8_scale = _p(res.h16), _p(res.q8), FMT_F16_Q8, res.q8_scale
        else:
            d.res_hi, d.res_lo, d.res_fmt = _p(res.hi), _p(res.lo), FMT_SPLIT_BF16
    d.out_f32, d.stats = _p(out.f32), _p(stats)
    if hq:
        d.out_hi, d.out_lo, d.out_fmt = _p(out.h16), _p(out.q8), FMT_F16_Q8
    else:
        d.out_hi, d.out_lo, d.out_fmt = _p(out.hi), _p(out.lo), FMT_SPLIT_BF16
    d.N, d.D, d.H, d.W, d.Cin, d.Cout = N, D, H, W, pw.Cin, pw.Cout
    d.KD, d.KH, d.KW = pw.k
    d.Cout_pad, d.gn_groups, d.act = pw.Cout_pad, stats_groups, act
    d.stride, d.in_C, d.out_C = 1, C, pw.Cout
    if pw.Cin2:
        d.Cin2, d.in2_C, d.in2_c_off, d.stride2 = pw.Cin2, src2.C, 0, 1
        d.in2_hi, d.in2_lo = _p(src2.h16), _p(src2.q8)
    L = _lib.load()
    flops = 2 * N * D * H * W * pw.Cout * (pw.Cin * pw.k[0] * pw.k[1] * pw.k[2] + pw.Cin2)
    with _Prof(("conv_tc_q8" if q8_in else "conv_tc") + f"|{N}x{D}x{H}x{W} {pw.Cin}->{pw.Cout} k{pw.k[0]}{pw.k[1]}{pw.k[2]} s1",
               flops):
        _lib.check(L.mp_conv_tc(ctypes.byref(d), _stream()), "mp_conv_tc")
    _count()
    return out, stats


_CONV_MODE = os.environ.get("MPB200_CONV", "auto")   # auto | tc | simt


def set_conv_mode(mode: str) -> None:
    global _CONV_MODE
    assert mode in ("auto", "tc", "simt")
    _CONV_MODE = mode


def conv(a: Act, pw: PackedConv, res: Optional[Act] = None, act: int = ACT_NONE, f32: bool = True,
         split: bool = False, stats_groups: int = 0, mode: Optional[str] = None, stride: int = 1, in_c_off: int = 0,
         out: Optional[Act] = None, out_c_off: int = 0, h16: bool = False, src2: Optional[Act] = None,
         stride2: int = 1, in2_c_off: int = 0, hq: bool = False,
         out_q8_scale: float = 1.0) -> Tuple[Act, Optional[torch.Tensor]]:
    """act(conv(a) + bias + res) -> (Act, GroupNorm statistics of the written values or None).

    `hq=True` writes the F16_Q8 plane pair (fp16 + FP8 bytes, the latter with the per-tensor power-of-two `out_q8_scale`).
    `src2` feeds the fused 1x1 shortcut of a `pack_conv(..., shortcut=...)` plan: an activation in the same operand
    format whose grid is `stride2` times the output grid.
    `pw.prec == PREC_F16X2` runs the two-pass fp16 kernel: `a` (and `res`, unless it is fp32) must carry an fp16 plane
    and `h16=True` asks for an fp16 output plane.

    `stride` (1|2) applies to H and W.  `in_c_off` selects the window [in_c_off, in_c_off + pw.Cin) of a's channels;
    `out` / `out_c_off` write into the channel window of an existing activation (grouped convolutions)."""
    if pw.prec == PREC_F16_Q8 or hq:
        return _conv_q8(a, pw, res, act, f32, split, stats_groups, hq, out_q8_scale, src2)
    half = pw.prec == PREC_F16X2
    if half:
        if a.h16 is None:
            raise RuntimeError("conv: a PREC_F16X2 convolution needs an fp16 activation plane")
        if split or stats_groups or (res is not None and res.f32 is None and res.h16 is None):
            raise RuntimeError("conv: PREC_F16X2 supports fp32 / fp16 outputs and residuals only, no GroupNorm statistics")
    elif h16:
        raise RuntimeError("conv: fp16 output planes are produced by PREC_F16X2 convolutions only")
    elif a.hi is None:
        ensure_split(a)
    N, D, H, W, C = a.shape
    if in_c_off + pw.Cin > C:
        raise RuntimeError(f"conv: activation has {C} channels, weights expect {pw.Cin} (+{in_c_off})")
    Ho, Wo = H // stride, W // stride
    if out is None:
        out = _alloc((N, D, Ho, Wo, pw.Cout), a.device, f32, split, h16)
    elif out.shape[:4] != (N, D, Ho, Wo) or out_c_off + pw.Cout > out.shape[4]:
        raise RuntimeError(f"conv: output window does not fit {out.shape}")
    # the tensor-core kernel fuses GroupNorm statistics only when a sample has >= 128 output positions; the tiny
    # FlowField volumes (4..32 positions per sample) get theirs from the stand-alone reduction instead
    late_stats = bool(stats_groups) and (D * Ho * Wo < 128) and (out.f32 is not None)
    stats = new_stats(N, stats_groups, a.device) if (stats_groups and not late_stats) else None
    d = ConvDesc()
    d.w_hi, d.w_lo, d.bias, d.prec, d.acc_scale = _p(pw.w_hi), _p(pw.w_lo), _p(pw.bias), pw.prec, pw.acc_scale
    d.in_hi, d.in_lo = (_p(a.h16), None) if half else (_p(a.hi), _p(a.lo))
    if res is not None:
        if res.shape != out.shape:
            raise RuntimeError(f"conv: residual shape {res.shape} != output shape {out.shape}")
        if res.f32 is not None:
            d.res_f32 = _p(res.f32)
        elif half:
            d.res_hi = _p(res.h16)
        else:
            d.res_hi, d.res_lo = _p(res.hi), _p(res.lo)
    d.out_f32, d.stats = _p(out.f32), _p(stats)
    d.out_hi, d.out_lo = (_p(out.h16), None) if half else (_p(out.hi), _p(out.lo))
    d.N, d.D, d.H, d.W, d.Cin, d.Cout = N, D, H, W, pw.Cin, pw.Cout
    d.KD, d.KH, d.KW = pw.k
    d.Cout_pad, d.gn_groups, d.act = pw.Cout_pad, (0 if late_stats else stats_groups), act
    d.stride, d.in_c_off, d.in_C, d.out_c_off, d.out_C = stride, in_c_off, C, out_c_off, out.shape[4]
    if pw.Cin2:
        if src2 is None or src2.shape[:4] != (N, D, Ho * stride2, Wo * stride2) or in2_c_off + pw.Cin2 > src2.shape[4]:
            raise RuntimeError(f"conv: the fused shortcut needs a second source on a {stride2}x grid with {pw.Cin2} channels")
        if half and src2.h16 is None:
            raise RuntimeError("conv: a PREC_F16X2 shortcut source needs an fp16 activation plane")
        if not half and src2.hi is None:
            ensure_split(src2)
        d.Cin2, d.in2_C, d.in2_c_off, d.stride2 = pw.Cin2, src2.shape[4], in2_c_off, stride2
        d.in2_hi, d.in2_lo = (_p(src2.h16), None) if half else (_p(src2.hi), _p(src2.lo))
    elif src2 is not None:
        raise RuntimeError("conv: src2 given but the packed weights hold no fused shortcut")
    L = _lib.load()
    mode = mode or _CONV_MODE
    use_tc = mode == "tc" or (mode == "auto" and L.mp_conv_tc_supported(ctypes.byref(d)) == 1)
    flops = 2 * N * D * Ho * Wo * pw.Cout * (pw.Cin * pw.k[0] * pw.k[1] * pw.k[2] + pw.Cin2)
    if half and not use_tc:
        raise RuntimeError(f"conv: shape {a.shape} x {pw.Cin}->{pw.Cout} is not supported by the fp16 two-pass kernel")
    with _Prof(("conv_tc" if use_tc else "conv_simt") + ("_h" if half else "") +
               f"|{N}x{D}x{H}x{W} {pw.Cin}->{pw.Cout} k{pw.k[0]}{pw.k[1]}{pw.k[2]} s{stride}" +
               (f" +sc{pw.Cin2}" if pw.Cin2 else ""), flops):
        if use_tc:
            _lib.check(L.mp_conv_tc(ctypes.byref(d), _stream()), "mp_conv_tc")
        else:
            _lib.check(L.mp_conv_simt(ctypes.byref(d), _stream()), "mp_conv_simt")
    _count()
    if late_stats:
        stats = gn_stats(out, stats_groups)
    return out, stats


def pack_conv_train(weight: torch.Tensor, bias: Optional[torch.Tensor], dgrad: bool = False) -> PackedConv:
    """`pack_conv(weight, bias)` (dgrad=False) or `pack_conv_dgrad(weight)` (dgrad=True) for the training step, where every weight
    is repacked twice per iteration: ONE kernel (`mp_pack_conv_weights`) instead of ~10 ATen launches; the planes are bit-identical
    to `pack_conv`'s for fp32 weights (split-bf16 only)."""
    w = weight.detach()
    if not w.is_cuda or w.dtype != torch.float32:
        return pack_conv_dgrad(w) if dgrad else pack_conv(w, bias)
    w = w.contiguous()
    Cout, Cin = w.shape[0], w.shape[1]
    k = tuple(w.shape[2:]) if w.dim() == 5 else (1,) + tuple(w.shape[2:])
    T = k[0] * k[1] * k[2]
    rows, cols = (Cin, Cout) if dgrad else (Cout, Cin)
    rows_pad = (rows + 15) // 16 * 16
    planes = torch.empty((2, rows_pad, T * cols), dtype=torch.bfloat16, device=w.device)
    L = _lib.load()
    _lib.check(L.mp_pack_conv_weights(_p(w), _p(planes[0]), _p(planes[1]), Cout, Cin, T, rows_pad, 1 if dgrad else 0, _stream()),
               "mp_pack_conv_weights")
    _count()
    b = None if (bias is None or dgrad) else bias.detach().float().contiguous()
    return PackedConv(planes[0], planes[1], b, cols, rows, rows_pad, k)


def pack_conv_dgrad(weight: torch.Tensor, device=None) -> PackedConv:
    """Weights of the DATA gradient of a stride-1 "same" convolution (row f-2): dX = conv(dY, W') with W'[ci, co, k] =
    W[co, ci, flip(k)] -- the transposed convolution is again an implicit GEMM of the same shape family, so it runs on
    the same tcgen05 kernel as the forward pass (three-pass split-bf16: fp32-grade gradients)."""
    w = weight.detach()
    dims = tuple(range(2, w.dim()))
    return pack_conv(w.transpose(0, 1).flip(dims).contiguous(), None, device or weight.device)


def conv_input_grad(grad_out: Act, pw_dgrad: PackedConv, f32: bool = True, split: bool = False) -> Act:
    """dL/dX of `conv(x, W)` (stride 1, padding k // 2) from dL/dY: `conv(grad_out, pack_conv_dgrad(W))`."""
    return conv(grad_out, pw_dgrad, f32=f32, split=split)[0]


@_profiled
def maxpool3x3s2(a: Act) -> Act:
    """nn.MaxPool2d(3, 2, 1) on a split 2-D activation."""
    N, D, H, W, C = a.shape
    assert D == 1
    ensure_split(a)
    out = _alloc((N, 1, H // 2, W // 2, C), a.device, False, True)
    L = _lib.load()
    _lib.check(L.mp_maxpool3x3s2_cl(_p(a.hi), _p(a.lo), _p(out.hi), _p(out.lo), N, H, W, C, _stream()),
               "mp_maxpool3x3s2_cl")
    _count()
    return out


@_profiled
def global_avgpool(a: Act) -> torch.Tensor:
    """nn.AdaptiveAvgPool2d(1) + flatten: -> [N, C] fp32."""
    N, D, H, W, C = a.shape
    out = torch.empty((N, C), dtype=torch.float32, device=a.device)
    L = _lib.load()
    _lib.check(L.mp_global_avgpool_cl(_p(a.f32), _p(a.hi), _p(a.lo), _p(out), N, D * H * W, C, _stream()),
               "mp_global_avgpool_cl")
    _count()
    return out


# ----------------------------------------------------------------------------------------------------- fp16 trunk ops
@_profiled
def im2col3x3_f16(x: torch.Tensor, stride: int = 1) -> Act:
    """NCHW fp32 RGB frames [N,3,H,W] -> fp16 CL patches [N,1,H/stride,W/stride,32] (see `pack_stem3x3_f16`)."""
    _chk_cuda(x, torch.float32, "im2col3x3_f16")
    N, C, H, W = x.shape
    out = _alloc((N, 1, H // stride, W // stride, 32), x.device, False, False, True)
    L = _lib.load()
    _lib.check(L.mp_im2col3x3_f16(_p(x), _p(out.h16), N, C, H, W, stride, _stream()), "mp_im2col3x3_f16")
    _count()
    return out


@_profiled
def maxpool3x3s2_f16(a: Act) -> Act:
    """nn.MaxPool2d(3, 2, 1) on an fp16 2-D activation."""
    N, D, H, W, C = a.shape
    assert D == 1 and a.h16 is not None
    out = _alloc((N, 1, H // 2, W // 2, C), a.device, False, False, True)
    L = _lib.load()
    _lib.check(L.mp_maxpool3x3s2_cl_f16(_p(a.h16), _p(out.h16), N, H, W, C, _stream()), "mp_maxpool3x3s2_cl_f16")
    _count()
    return out


@_profiled
def stem3x3_relu_maxpool_f16(x: torch.Tensor, pw: PackedConv) -> Act:
    """conv3x3(3 -> C) + folded BN + ReLU + MaxPool2d(3, 2, 1) in one kernel (resnet.py:192-197, 271-273): NCHW fp32 frames
    [N,3,H,W] -> fp16 CL [N,1,H/2,W/2,C]; `pw` = `pack_stem3x3_f16(...)`."""
    _chk_cuda(x, torch.float32, "stem3x3_relu_maxpool_f16")
    N, C, H, W = x.shape
    if C != 3 or pw.prec != PREC_F16X2 or pw.Cin != 32 or pw.Cout != pw.Cout_pad:
        raise RuntimeError(f"stem3x3_relu_maxpool_f16: needs RGB frames and a pack_stem3x3_f16 weight pack, got {tuple(x.shape)}")
    out = _alloc((N, 1, H // 2, W // 2, pw.Cout), x.device, False, False, True)
    L = _lib.load()
    _lib.check(L.mp_stem3x3_relu_maxpool_f16(_p(x), _p(pw.w_hi), _p(pw.w_lo), _p(pw.bias) if pw.bias is not None else None,
                                             float(pw.acc_scale), _p(out.h16), N, H, W, pw.Cout, _stream()),
               "mp_stem3x3_relu_maxpool_f16")
    _count()
    return out


@_profiled
def global_avgpool_f16(a: Act) -> torch.Tensor:
    """nn.AdaptiveAvgPool2d(1) + flatten on an fp16 activation: -> [N, C] fp32."""
    N, D, H, W, C = a.shape
    out = torch.empty((N, C), dtype=torch.float32, device=a.device)
    L = _lib.load()
    _lib.check(L.mp_global_avgpool_cl_f16(_p(a.h16), _p(out), N, D * H * W, C, _stream()), "mp_global_avgpool_cl_f16")
    _count()
    return out


# ----------------------------------------------------------------------------------------------------- warping
def _brick_ok(W: int, Wo: int, D: int, H: int, *tensors) -> bool:
    """Shapes the brick-staged kernels accept (16-byte TMA rows, 10-bit packed cells, aligned bases)."""
    return W % 4 == 0 and Wo % 4 == 0 and max(D, H, W) <= 1023 and all(t.data_ptr() % 16 == 0 for t in tensors)


def gs_brick_tune(cfg: Sequence[int] = ()) -> None:
    """Profiling hook: {tz, ty, tx, BD, BH, BW, threads, groups} overrides of the brick kernels (0 / () = automatic)."""
    arr = (ctypes.c_int * 8)(*([int(c) for c in cfg] + [0] * (8 - len(cfg))))
    _lib.check(_lib.load().mp_gs_brick_tune(arr, 8), "mp_gs_brick_tune")


def _brick_ws(L, N, Do, Ho, Wo, device, second_pass: bool):
    if not second_pass:
        return None, 0
    nb = L.mp_gs_brick_workspace_bytes(N, Do, Ho, Wo)
    return torch.empty(nb, dtype=torch.uint8, device=device), nb


def grid_sample3d(v: torch.Tensor, grid: torch.Tensor, direct: bool = False, impl: Optional[str] = None,
                  bucket: bool = True, second_pass: bool = True) -> torch.Tensor:
    """F.grid_sample(v, grid, 'bilinear', 'border', align_corners=True) for NCDHW fp32 (model.py:1062).

    `impl`: "brick" (default where the shape allows: TMA-staged bricks in shared memory; tiles whose sampling region does
    not fit a brick are gathered by a second launch, or inside the brick kernel with `second_pass=False`), "ws"
    (channels-last workspace copy + gather: the better choice for random-permutation grids), "direct" (plain NCDHW
    gather).  All three give the same bits."""
    _chk_cuda(v, torch.float32, "grid_sample3d v")
    _chk_cuda(grid, torch.float32, "grid_sample3d grid")
    N, C, D, H, W = v.shape
    Ng, Do, Ho, Wo, three = grid.shape
    if Ng != N or three != 3:
        raise RuntimeError("grid_sample3d: grid must be [N, Do, Ho, Wo, 3]")
    out = torch.empty((N, C, Do, Ho, Wo), dtype=torch.float32, device=v.device)
    L = _lib.load()
    if impl is None:
        impl = "direct" if direct else ("brick" if _brick_ok(W, Wo, D, H, v, out) else "ws")
    nbytes = N * (C * D * H * W + C * Do * Ho * Wo + 3 * Do * Ho * Wo) * 4
    if impl == "brick":
        ws, nb = _brick_ws(L, N, Do, Ho, Wo, v.device, second_pass)
        with _Prof("grid_sample3d", 0, nbytes):
            _lib.check(L.mp_grid_sample3d_brick(_p(v), _p(grid), _p(out), _p(ws), nb, N, C, D, H, W, Do, Ho, Wo,
                                                0 if bucket else 1, _stream()), "mp_grid_sample3d_brick")
        _count(2 if second_pass else 1)
        return out
    ws_bytes = L.mp_gather_workspace_bytes(N, C, D, H, W) if impl == "ws" else 0
    if ws_bytes:
        ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=v.device)
        with _Prof("grid_sample3d", 0, nbytes):
            _lib.check(L.mp_grid_sample3d_ws(_p(v), _p(grid), _p(out), _p(ws), ws_bytes, N, C, D, H, W, Do, Ho, Wo,
                                             _stream()), "mp_grid_sample3d_ws")
        _count(2)
    else:
        _lib.check(L.mp_grid_sample3d(_p(v), _p(grid), _p(out), N, C, D, H, W, Do, Ho, Wo, _stream()), "mp_grid_sample3d")
        _count()
    return out


def apply_warping_field_ncdhw(v: torch.Tensor, warp_field: torch.Tensor, direct: bool = False,
                              impl: Optional[str] = None) -> torch.Tensor:
    """apply_warping_field(v, warp_field) (model.py:1028-1065), NCDHW fp32; `impl` as in `grid_sample3d`."""
    _chk_cuda(v, torch.float32, "apply_warping_field v")
    _chk_cuda(warp_field, torch.float32, "apply_warping_field warp_field")
    N, C, D, H, W = v.shape
    Nf, three, Df, Hf, Wf = warp_field.shape
    if Nf != N or three != 3:
        raise RuntimeError("apply_warping_field: warp_field must be [N, 3, Df, Hf, Wf]")
    out = torch.empty_like(v)
    L = _lib.load()
    if impl is None:
        impl = "direct" if direct else ("brick" if (_brick_ok(W, W, D, H, v, out) and min(D, H, W) > 1) else "ws")
    if impl == "brick":
        ws, nb = _brick_ws(L, N, D, H, W, v.device, True)
        _lib.check(L.mp_apply_warping_field_brick(_p(v), _p(warp_field), _p(out), _p(ws), nb, N, C, D, H, W, Df, Hf, Wf, 0,
                                                  _stream()), "mp_apply_warping_field_brick")
        _count(2)
        return out
    ws_bytes = L.mp_gather_workspace_bytes(N, C, D, H, W) if impl == "ws" else 0
    if ws_bytes:
        ws = torch.empty(ws_bytes // 4, dtype=torch.float32, device=v.device)
        _lib.check(L.mp_apply_warping_field_ws(_p(v), _p(warp_field), _p(out), _p(ws), ws_bytes, N, C, D, H, W, Df, Hf,
                                               Wf, _stream()), "mp_apply_warping_field_ws")
        _count(2)
    else:
        _lib.check(L.mp_apply_warping_field(_p(v), _p(warp_field), _p(out), N, C, D, H, W, Df, Hf, Wf, _stream()),
                   "mp_apply_warping_field")
        _count()
    return out


def apply_warping_field_backward(grad_out: torch.Tensor, v: torch.Tensor, warp_field: torch.Tensor, need_v: bool = True,
                                 need_wf: bool = True):
    """Gradients of `apply_warping_field_ncdhw` w.r.t. v and warp_field (row f-2; ATen grid_sampler_3d_backward semantics)."""
    for t, nm in ((grad_out, "grad_out"), (v, "v"), (warp_field, "warp_field")):
        _chk_cuda(t, torch.float32, "apply_warping_field_backward " + nm)
    N, C, D, H, W = v.shape
    _, _, Df, Hf, Wf = warp_field.shape
    gv = torch.zeros_like(v) if need_v else None
    gwf = torch.zeros_like(warp_field) if need_wf else None
    L = _lib.load()
    _lib.check(L.mp_apply_warping_field_backward(_p(grad_out), _p(v), _p(warp_field), _p(gv), _p(gwf), N, C, D, H, W, Df, Hf, Wf,
                                                 _stream()), "mp_apply_warping_field_backward")
    _count()
    return gv, gwf


class WarpFunction(torch.autograd.Function):
    """`apply_warping_field(v, warp_field)` with a CUDA backward (row f-2): forward = the brick-staged kernel, backward =
    `mp_apply_warping_field_backward`."""

    @staticmethod
    def forward(ctx, v, warp_field):
        v, warp_field = v.detach().float().contiguous(), warp_field.detach().float().contiguous()
        ctx.save_for_backward(v, warp_field)
        return apply_warping_field_ncdhw(v, warp_field)

    @staticmethod
    def backward(ctx, grad_out):
        v, warp_field = ctx.saved_tensors
        gv, gwf = apply_warping_field_backward(grad_out.float().contiguous(), v, warp_field, ctx.needs_input_grad[0],
                                               ctx.needs_input_grad[1])
        return gv, gwf


# weight gradient on the tcgen05 kernel (MPB200_WGRAD=mma selects the first-generation mma.sync kernel; A/B runs and tests)
WGRAD_TC = os.environ.get("MPB200_WGRAD", "tc") != "mma"


def conv_weight_grad(x: Act, grad_out: Act, k: Tuple[int, int, int]) -> torch.Tensor:
    """dL/dW of a stride-1 "same" convolution (row f-2): x, grad_out channels-last fp32 Acts -> (Cout, Cin, kd, kh, kw) fp32.
    One tensor-core GEMM per filter tap with K = positions (three bf16 passes, fp32 accumulation)."""
    N, D, H, W, Cin = x.shape
    Cout = grad_out.shape[-1]
    if tuple(grad_out.shape[:4]) != (N, D, H, W) or (x.f32 is None and x.hi is None) or (grad_out.f32 is None and grad_out.hi is None):
        raise RuntimeError(f"conv_weight_grad: channels-last tensors of one spatial size, got {x.shape} / {grad_out.shape}")
    kd, kh, kw = k
    dw = torch.empty((Cout, kd * kh * kw, Cin), dtype=torch.float32, device=x.device)
    L = _lib.load()
    if WGRAD_TC and L.mp_conv_wgrad_tc_supported(Cin, Cout, kd, kh, kw) == 1:
        # tcgen05 path: both operands as split-bf16 planes, read position-major straight from the channels-last tensors
        ensure_split(x)
        ensure_split(grad_out)
        flops = 2 * N * D * H * W * Cout * Cin * kd * kh * kw
        with _Prof(f"wgrad_tc|{N}x{D}x{H}x{W} {Cin}->{Cout} k{kd}{kh}{kw}", flops):
            _lib.check(L.mp_conv_wgrad_tc(_p(x.hi), _p(x.lo), _p(grad_out.hi), _p(grad_out.lo), _p(dw), N, D, H, W, Cin, Cout,
                                          kd, kh, kw, _stream()), "mp_conv_wgrad_tc")
    else:
        for t in (x, grad_out):                  # the mma.sync kernel reads fp32: rebuild it from the planes if need be
            if t.f32 is None:
                t.f32 = t.hi.float() + t.lo.float()
        _lib.check(L.mp_conv_wgrad(_p(x.f32), _p(grad_out.f32), _p(dw), N, D, H, W, Cin, Cout, kd, kh, kw, _stream()),
                   "mp_conv_wgrad")
    _count(2)
    return dw.view(Cout, kd, kh, kw, Cin).permute(0, 4, 1, 2, 3).contiguous()


def bias_grad(grad_out: Act) -> torch.Tensor:
    C = grad_out.shape[-1]
    db = torch.empty((C,), dtype=torch.float32, device=grad_out.device)
    L = _lib.load()
    _lib.check(L.mp_bias_grad(_p(grad_out.f32), _p(db), grad_out.f32.numel() // C, C, _stream()), "mp_bias_grad")
    _count(2)
    return db


def group_norm_backward(x: Act, grad_out: Act, stats: torch.Tensor, G: int, gamma: Optional[torch.Tensor], eps: float = 1e-5):
    """nn.GroupNorm backward on channels-last fp32 Acts: -> (dx Act, dgamma [C] fp32, dbeta [C] fp32)."""
    N, D, H, W, C = x.shape
    dx = _alloc(x.shape, x.device, True, False)
    dg = torch.empty((C,), dtype=torch.float64, device=x.device)
    db = torch.empty((C,), dtype=torch.float64, device=x.device)
    ws = torch.empty((N * (2 * G + 2 * C),), dtype=torch.float64, device=x.device)     # group sums + coefficient table
    L = _lib.load()
    _lib.check(L.mp_group_norm_backward(_p(x.f32), _p(grad_out.f32), _p(stats), _p(gamma), _p(dx.f32), _p(dg), _p(db), _p(ws),
                                        N, D * H * W, C, G, eps, _stream()), "mp_group_norm_backward")
    _count(5)
    return dx, dg.float(), db.float()


def _to_cl_act(t: torch.Tensor) -> Act:
    """Logical NCHW / NCDHW fp32 tensor -> channels-last fp32 Act.  The autograd Functions hand each other logical-NCHW tensors
    whose MEMORY is already channels-last (`_cl_view`): those are wrapped without a copy; a contiguous NCHW tensor (the reference
    layout, at the boundary) goes through the transposing kernel; anything else through one strided ATen copy."""
    t = t.detach()
    if t.dtype != torch.float32:
        t = t.float()
    if t.dim() == 4:
        v = t.permute(0, 2, 3, 1).unsqueeze(1)
    else:
        v = t.permute(0, 2, 3, 4, 1)
    if not v.is_contiguous():
        if t.is_contiguous():
            return from_nchw(t, f32=True, split=False)
        v = v.contiguous()
    return Act(tuple(v.shape), f32=v)


def _cl_view(a: Act, ndim: int) -> torch.Tensor:
    """Channels-last fp32 Act -> logical NCHW / NCDHW tensor that is a VIEW of the same memory (torch.channels_last /
    channels_last_3d strides): no transposition between consecutive operators of the differentiable path."""
    N, D, H, W, C = a.shape
    v = a.f32.view(N, D, H, W, C)
    return v.permute(0, 4, 1, 2, 3) if ndim == 5 else v.view(N, H, W, C).permute(0, 3, 1, 2)


class ConvFunction(torch.autograd.Function):
    """`F.conv2d` / `F.conv3d` (stride 1, padding k // 2) on libmpb200 with CUDA backward (row f-2): forward = the tcgen05
    implicit GEMM, dX = the same kernel on flipped / transposed weights, dW = `mp_conv_wgrad`, db = `mp_bias_grad`.
    Tensors are logical NCHW / NCDHW fp32; outputs (and gradients) are channels-last in memory (`_cl_view`), so a chain of these
    Functions never transposes; the reference layout is restored by `.contiguous()` at the modules' public `forward()`."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        nd = x.dim()
        ctx.nd = nd
        a = _to_cl_act(x)
        ensure_split(a)
        k = tuple(weight.shape[2:]) if nd == 5 else (1,) + tuple(weight.shape[2:])
        ctx.k = k
        ctx.save_for_backward(x.detach(), weight.detach())
        ctx.has_bias = bias is not None
        out, _ = conv(a, pack_conv_train(weight, bias), f32=True)
        return _cl_view(out, nd)

    @staticmethod
    def backward(ctx, grad_out):
        x, weight = ctx.saved_tensors
        g = _to_cl_act(grad_out)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            ensure_split(g)
            gx = _cl_view(conv_input_grad(g, pack_conv_train(weight, None, dgrad=True)), ctx.nd)
        if ctx.needs_input_grad[1]:
            gw = conv_weight_grad(_to_cl_act(x), g, ctx.k)
            if ctx.nd == 4:
                gw = gw.squeeze(2)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = bias_grad(g)
        return gx, gw, gb


class ConvStatsFunction(torch.autograd.Function):
    """`ConvFunction` whose forward also returns the per-(sample, group) sum / sum of squares of its output, accumulated by the
    convolution's epilogue (`mp_conv_desc.stats`) -- the normalisation that follows (`GroupNormFunction` / `BatchNormFunction` with
    `stats=`) then needs no statistics pass over the tensor.  `groups` = GroupNorm groups, or the channel count for BatchNorm."""

    @staticmethod
    def forward(ctx, x, weight, bias, groups, stride):
        nd = x.dim()
        ctx.nd, ctx.stride = nd, stride
        a = _to_cl_act(x)
        ensure_split(a)
        ctx.k = tuple(weight.shape[2:]) if nd == 5 else (1,) + tuple(weight.shape[2:])
        ctx.save_for_backward(x.detach(), weight.detach())
        ctx.has_bias = bias is not None
        out, stats = conv(a, pack_conv_train(weight, bias), f32=True, stats_groups=groups, stride=stride)
        ctx.mark_non_differentiable(stats)
        return _cl_view(out, nd), stats

    @staticmethod
    def backward(ctx, grad_out, _gstats):
        fn = ConvS2Function if ctx.stride == 2 else ConvFunction
        return fn.backward(ctx, grad_out) + (None, None)


class GroupNormFunction(torch.autograd.Function):
    """`F.group_norm` on libmpb200 with CUDA backward (row f-2); NCHW / NCDHW fp32 in and out."""

    @staticmethod
    def forward(ctx, x, G, gamma, beta, eps, stats=None):
        a = _to_cl_act(x)
        if stats is None:
            stats = gn_stats(a, G)
        ab = gn_finalize(stats, a.shape, G, None if gamma is None else gamma.detach().float().contiguous(),
                         None if beta is None else beta.detach().float().contiguous(), eps=eps)
        ctx.save_for_backward(x.detach(), stats, None if gamma is None else gamma.detach().float().contiguous())
        ctx.G, ctx.eps, ctx.nd = G, eps, x.dim()
        return _cl_view(affine_act(a, ab, None, ACT_NONE, f32=True, split=False), x.dim())

    @staticmethod
    def backward(ctx, grad_out):
        x, stats, gamma = ctx.saved_tensors
        dx, dg, db = group_norm_backward(_to_cl_act(x), _to_cl_act(grad_out), stats, ctx.G, gamma, ctx.eps)
        return (_cl_view(dx, ctx.nd), None, dg if ctx.needs_input_grad[2] else None,
                db if ctx.needs_input_grad[3] else None, None, None)


class ConvS2Function(torch.autograd.Function):
    """`F.conv2d(x, w, b, stride=2, padding=k // 2)` (odd k, even H and W) with CUDA forward and backward (row f-2).  Forward = the
    tcgen05 kernel's native stride-2 walk (TMA element strides).  A stride-2 "same" convolution is the stride-1 one sampled at the
    even positions, so its gradients are the stride-1 gradients of dY spread onto the even positions of a zero tensor: dX and dW
    reuse `conv_input_grad` / `conv_weight_grad` unchanged (4x the minimal work on the few stride-2 layers of the encoders)."""

    @staticmethod
    def forward(ctx, x, weight, bias):
        if x.dim() != 4 or x.shape[2] % 2 or x.shape[3] % 2:
            raise RuntimeError(f"ConvS2Function: NCHW input with even H and W expected, got {tuple(x.shape)}")
        a = _to_cl_act(x)
        ensure_split(a)
        ctx.k = (1,) + tuple(weight.shape[2:])
        ctx.save_for_backward(x.detach(), weight.detach())
        ctx.has_bias = bias is not None
        out, _ = conv(a, pack_conv_train(weight, bias), f32=True, stride=2)
        return _cl_view(out, 4)

    @staticmethod
    def backward(ctx, grad_out):
        x, weight = ctx.saved_tensors
        N, _, H, W = x.shape
        up = torch.zeros((N, H, W, grad_out.shape[1]), dtype=torch.float32, device=grad_out.device).permute(0, 3, 1, 2)
        up[:, :, ::2, ::2] = grad_out
        g = _to_cl_act(up)
        gx = gw = gb = None
        if ctx.needs_input_grad[0]:
            ensure_split(g)
            gx = _cl_view(conv_input_grad(g, pack_conv_train(weight, None, dgrad=True)), 4)
        if ctx.needs_input_grad[1]:
            gw = conv_weight_grad(_to_cl_act(x), g, ctx.k).squeeze(2)
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = bias_grad(g)
        return gx, gw, gb


def upsample2x_linear_backward(grad_out: Act, up_d: int) -> Act:
    """Adjoint of `upsample2x_linear` on channels-last fp32 Acts (row f-2)."""
    N, Do, Ho, Wo, C = grad_out.shape
    out = _alloc((N, Do // up_d, Ho // 2, Wo // 2, C), grad_out.device, True, False)
    L = _lib.load()
    _lib.check(L.mp_upsample2x_linear_backward_cl(_p(grad_out.f32), _p(out.f32), N, Do // up_d, Ho // 2, Wo // 2, C, up_d,
                                                  _stream()), "mp_upsample2x_linear_backward_cl")
    _count()
    return out


class AvgPool2Function(torch.autograd.Function):
    """`F.avg_pool2d(x, 2)` (pool_d = 1) / `F.avg_pool3d(x, 2)` (pool_d = 2) on libmpb200 (row f-2): forward `mp_avgpool2_cl`,
    backward = the nearest upsample of the scaled gradient (`mp_upsample_nearest_cl`)."""

    @staticmethod
    def forward(ctx, x, pool_d):
        ctx.pool_d, ctx.nd = pool_d, x.dim()
        return _cl_view(avgpool2(_to_cl_act(x), pool_d, f32=True, split=False), x.dim())

    @staticmethod
    def backward(ctx, grad_out):
        g = _to_cl_act(grad_out * (1.0 / (4 * ctx.pool_d)))
        return _cl_view(upsample_nearest(g, (ctx.pool_d, 2, 2), f32=True, split=False), ctx.nd), None


class UpsampleNearestFunction(torch.autograd.Function):
    """`F.interpolate(x, scale_factor=(sd, 2, 2), mode="nearest")`, sd in {1, 2} (FlowField, model.py:427-433) on libmpb200:
    forward `mp_upsample_nearest_cl`, backward = the 2 x 2 (x sd) sum pool (`mp_avgpool2_cl` times the window size)."""

    @staticmethod
    def forward(ctx, x, sd):
        ctx.sd, ctx.nd = sd, x.dim()
        return _cl_view(upsample_nearest(_to_cl_act(x), (sd, 2, 2), f32=True, split=False), x.dim())

    @staticmethod
    def backward(ctx, grad_out):
        g = avgpool2(_to_cl_act(grad_out), ctx.sd, f32=True, split=False)
        return _cl_view(g, ctx.nd) * float(4 * ctx.sd), None


class UpsampleLinear2xFunction(torch.autograd.Function):
    """`nn.Upsample(scale_factor=2, mode="bilinear" | "trilinear", align_corners=True)` (G2d / G3d, model.py:585-589, 733-743) on
    libmpb200: forward `mp_upsample2x_linear_cl`, backward `mp_upsample2x_linear_backward_cl` (deterministic gather)."""

    @staticmethod
    def forward(ctx, x):
        ctx.nd = x.dim()
        ctx.up_d = 2 if x.dim() == 5 else 1
        return _cl_view(upsample2x_linear(_to_cl_act(x), ctx.up_d, f32=True, split=False), x.dim())

    @staticmethod
    def backward(ctx, grad_out):
        return _cl_view(upsample2x_linear_backward(_to_cl_act(grad_out), ctx.up_d), ctx.nd)


def im2col_rgb_split(x: torch.Tensor, kh: int, kw: int, stride: int) -> Act:
    """NCHW fp32 [N, C <= 4, H, W] -> split-bf16 patch rows [N, 1, H/stride, W/stride, Kpad] (Kpad = kh*kw*C rounded up to 16)."""
    _chk_cuda(x, torch.float32, "im2col_rgb_split")
    N, C, H, W = x.shape
    kpad = (kh * kw * C + 15) // 16 * 16
    out = _alloc((N, 1, H // stride, W // stride, kpad), x.device, False, True)
    L = _lib.load()
    _lib.check(L.mp_im2col_rgb_split(_p(x), _p(out.hi), _p(out.lo), N, C, H, W, kh, kw, stride, kpad, _stream()),
               "mp_im2col_rgb_split")
    _count()
    return out


class RgbStemConvFunction(torch.autograd.Function):
    """The RGB stem convolutions (Eapp 7x7, model.py:211; the stems of the three ResNets): `F.conv2d(x, w, b, stride, padding=k//2)`
    with x [N, 3, H, W].  Forward: the tcgen05 kernel with the three channels zero-padded to 16 (as in inference).  Weight gradient:
    instead of kh*kw tap GEMMs over 16 mostly-zero channels, the frame is unfolded into patch rows (`mp_im2col_rgb_split`,
    kh*kw*3 -> Kpad columns) and dW is ONE K = positions GEMM on tcgen05 (`mp_conv_wgrad_tc`, 1x1 filter, Cin = Kpad) at the OUTPUT
    resolution (weight-gradient total of a training iteration: 16.2 -> 12.9 ms); a stride-2 stem needs no zero-spread gradient.  The data gradient (only asked for
    when the frame itself requires grad: `Gbase.motionEncoder(generated_frame)`, train.py:289) takes the generic path.
    `groups` > 0 also returns the epilogue's normalisation statistics."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, groups):
        x = x.detach().float().contiguous()
        cout, cin, kh, kw = weight.shape
        a = from_nchw_pad16(x)
        wpad = _pad_channels(weight.detach().float(), 1, 16)
        ctx.save_for_backward(x, weight.detach())
        ctx.stride, ctx.has_bias = stride, bias is not None
        out, stats = conv(a, pack_conv_train(wpad, bias), f32=True, stride=stride, stats_groups=groups)
        if stats is None:
            stats = torch.empty(0, dtype=torch.float64, device=x.device)
        ctx.mark_non_differentiable(stats)
        return _cl_view(out, 4), stats

    @staticmethod
    def backward(ctx, grad_out, _gstats):
        x, weight = ctx.saved_tensors
        cout, cin, kh, kw = weight.shape
        g = _to_cl_act(grad_out)
        gx = gw = gb = None
        if ctx.needs_input_grad[1]:
            cols = im2col_rgb_split(x, kh, kw, ctx.stride)
            dw = conv_weight_grad(cols, g, (1, 1, 1))                                   # [Cout, Kpad, 1, 1, 1]
            gw = dw.reshape(cout, -1)[:, :kh * kw * cin].reshape(cout, kh, kw, cin).permute(0, 3, 1, 2).contiguous()
        if ctx.has_bias and ctx.needs_input_grad[2]:
            gb = bias_grad(g)
        if ctx.needs_input_grad[0]:
            N, _, H, W = x.shape
            if ctx.stride == 2:
                up = torch.zeros((N, H, W, cout), dtype=torch.float32, device=x.device).permute(0, 3, 1, 2)
                up[:, :, ::2, ::2] = grad_out
                g = _to_cl_act(up)
            ensure_split(g)
            wpad = _pad_channels(weight.float(), 1, 16)
            gx = _cl_view(conv_input_grad(g, pack_conv_train(wpad, None, dgrad=True)), 4)[:, :cin]
        return gx, gw, gb, None, None


def maxpool3x3s2_forward_idx(a: Act):
    """nn.MaxPool2d(3, 2, 1) on a channels-last fp32 Act -> (Act, uint8 window positions of the maxima)."""
    N, D, H, W, C = a.shape
    out = _alloc((N, 1, H // 2, W // 2, C), a.device, True, False)
    idx = torch.empty((N, 1, H // 2, W // 2, C), dtype=torch.uint8, device=a.device)
    L = _lib.load()
    _lib.check(L.mp_maxpool3x3s2_forward_idx(_p(a.f32), _p(out.f32), _p(idx), N, H, W, C, _stream()), "mp_maxpool3x3s2_forward_idx")
    _count()
    return out, idx


def maxpool3x3s2_backward(grad_out: Act, idx: torch.Tensor) -> Act:
    N, _, Ho, Wo, C = grad_out.shape
    out = _alloc((N, 1, Ho * 2, Wo * 2, C), grad_out.device, True, False)
    L = _lib.load()
    _lib.check(L.mp_maxpool3x3s2_backward(_p(grad_out.f32), _p(idx), _p(out.f32), N, Ho * 2, Wo * 2, C, _stream()),
               "mp_maxpool3x3s2_backward")
    _count()
    return out


class MaxPool3x3s2Function(torch.autograd.Function):
    """`nn.MaxPool2d(kernel_size=3, stride=2, padding=1)` (the ResNet stems) on libmpb200 with CUDA backward (row f-2): the forward
    stores one byte per output element (which of the 9 window positions held the first maximum, ATen's rule), the backward is a
    deterministic gather through them."""

    @staticmethod
    def forward(ctx, x):
        out, idx = maxpool3x3s2_forward_idx(_to_cl_act(x))
        ctx.save_for_backward(idx)
        return _cl_view(out, 4)

    @staticmethod
    def backward(ctx, grad_out):
        (idx,) = ctx.saved_tensors
        return _cl_view(maxpool3x3s2_backward(_to_cl_act(grad_out), idx), 4)


def _pad_channels(t: torch.Tensor, dim: int, mult: int) -> torch.Tensor:
    c = t.shape[dim]
    extra = (-c) % mult
    if not extra:
        return t
    shape = list(t.shape)
    shape[dim] = extra
    return torch.cat((t, t.new_zeros(shape)), dim)


def conv_train(x: torch.Tensor, weight: torch.Tensor, bias: Optional[torch.Tensor] = None, stride: int = 1,
               stats_groups: int = 0):
    """Differentiable `F.conv2d / F.conv3d(x, weight, bias, stride, padding=k // 2)` on libmpb200 (row f-2) for every convolution of
    the path: channel counts that are not multiples of 16 (RGB stems, the 3-channel heads) are zero-padded around the Function by
    torch expressions (autograd slices the gradients back), 1x1 stride-2 shortcuts sub-sample their input first, other stride-2
    convolutions (2-D only) go through `ConvS2Function`."""
    cin, cout = weight.shape[1], weight.shape[0]
    if x.shape[1] != cin:
        raise RuntimeError(f"conv_train: input has {x.shape[1]} channels, weight expects {cin} (grouped convolutions are not on "
                           "the trained path)")
    if stride not in (1, 2):
        raise RuntimeError(f"conv_train: stride {stride} is not supported")
    x = x.float()
    if x.dim() == 4 and cin <= 4 and cout % 16 == 0 and WGRAD_TC and all(k % 2 == 1 for k in weight.shape[2:]) \
            and x.shape[2] % stride == 0 and x.shape[3] % stride == 0:
        y, stats = RgbStemConvFunction.apply(x, weight, bias, stride, stats_groups)      # RGB stem: im2col weight gradient
        return (y, stats) if stats_groups else y
    if stride == 2 and all(k == 1 for k in weight.shape[2:]):
        x = x[:, :, ::2, ::2] if x.dim() == 4 else x[:, :, ::2, ::2, ::2]
        stride = 1
    x = _pad_channels(x, 1, 16)
    w = _pad_channels(_pad_channels(weight.float(), 1, 16), 0, 16)
    b = None if bias is None else _pad_channels(bias.float(), 0, 16)
    if stats_groups:
        # statistics of the output from the convolution's epilogue: -> (y, stats [N, groups, 2] float64).  Only when the groups
        # survive the channel padding (no padding, or one group per channel: padded channels are exactly zero)
        if w.shape[0] == cout or stats_groups == cout:
            y, stats = ConvStatsFunction.apply(x, w, b, stats_groups if w.shape[0] == cout else w.shape[0], stride)
            return (y, stats) if cout == y.shape[1] else (y[:, :cout], stats[:, :cout])
        y = (ConvS2Function if stride == 2 else ConvFunction).apply(x, w, b)
        return y[:, :cout], None
    y = (ConvS2Function if stride == 2 else ConvFunction).apply(x, w, b)
    return y if cout == y.shape[1] else y[:, :cout]


def conv_bn_train(x: torch.Tensor, conv_mod, bn, stride: Optional[int] = None) -> torch.Tensor:
    """`bn(conv(x))` of the differentiable path: in train mode the BatchNorm's batch statistics come out of the convolution's
    epilogue (one group per channel, summed over the samples), so the normalisation costs one read + one write of the tensor."""
    st = conv_mod.stride[0] if stride is None else stride
    if bn.training and not (st == 2 and all(k == 1 for k in conv_mod.weight.shape[2:])):
        y, stats = conv_train(x, conv_mod.weight, conv_mod.bias, stride=st, stats_groups=conv_mod.weight.shape[0])
        return batch_norm_train(y, bn, None if stats is None else stats.sum(dim=0, keepdim=True))
    return batch_norm_train(conv_train(x, conv_mod.weight, conv_mod.bias, stride=st), bn)


class BatchNormFunction(torch.autograd.Function):
    """Train-mode `F.batch_norm` (batch statistics) with CUDA forward and backward (row f-2).  On the channels-last tensor a
    BatchNorm over (N, H, W) is a GroupNorm with ONE sample of N*H*W rows and one group per channel, so forward and backward are
    the GroupNorm entry points called with N = 1, G = C (`mp_gn_stats`, `mp_gn_finalize`, `mp_affine_act_cl`,
    `mp_group_norm_backward`: double-precision sums).  Returns (y, batch mean, biased batch variance); the caller folds the last two
    into the running statistics."""

    @staticmethod
    def forward(ctx, x, gamma, beta, eps, stats=None):
        a = _to_cl_act(x)
        N, D, H, W, C = a.shape
        a1 = Act((1, 1, N * D * H, W, C), f32=a.f32.view(1, 1, N * D * H, W, C))
        if stats is None:
            stats = gn_stats(a1, C)
        g = None if gamma is None else gamma.detach().float().contiguous()
        ab = gn_finalize(stats, a1.shape, C, g, None if beta is None else beta.detach().float().contiguous(), eps=eps)
        ctx.save_for_backward(x.detach(), stats, g)
        ctx.eps, ctx.nd = eps, x.dim()
        y = affine_act(a1, ab, None, ACT_NONE, f32=True, split=False)
        cnt = float(N * D * H * W)
        mean = stats[0, :, 0] / cnt
        var = (stats[0, :, 1] / cnt - mean * mean).clamp_min(0.0)
        ctx.mark_non_differentiable(mean, var)
        return _cl_view(Act(a.shape, f32=y.f32.view(a.shape)), x.dim()), mean, var

    @staticmethod
    def backward(ctx, grad_out, _gm, _gv):
        x, stats, gamma = ctx.saved_tensors
        a, g = _to_cl_act(x), _to_cl_act(grad_out)
        N, D, H, W, C = a.shape
        one = (1, 1, N * D * H, W, C)
        dx, dg, db = group_norm_backward(Act(one, f32=a.f32.view(one)), Act(one, f32=g.f32.view(one)), stats, C, gamma, ctx.eps)
        return (_cl_view(Act(a.shape, f32=dx.f32.view(a.shape)), ctx.nd), dg if ctx.needs_input_grad[1] else None,
                db if ctx.needs_input_grad[2] else None, None, None)


def batch_norm_train(x: torch.Tensor, bn, stats: Optional[torch.Tensor] = None) -> torch.Tensor:
    """`nn.BatchNorm2d.forward` for the differentiable path (row f-2).  Train mode: batch statistics through `BatchNormFunction`
    and the running statistics updated like ATen does (momentum, unbiased variance, `num_batches_tracked`); eval mode: the
    running-statistics affine map as a torch expression (elementwise)."""
    if bn.training or bn.running_mean is None:
        y, mean, var = BatchNormFunction.apply(x.float(), bn.weight, bn.bias, bn.eps, stats)
        if bn.training and bn.track_running_stats and bn.running_mean is not None:
            with torch.no_grad():
                n = x.numel() // x.shape[1]
                bn.num_batches_tracked += 1
                m = bn.momentum if bn.momentum is not None else 1.0 / float(bn.num_batches_tracked)
                bn.running_mean.mul_(1 - m).add_(mean.to(bn.running_mean.dtype), alpha=m)
                bn.running_var.mul_(1 - m).add_((var * (n / max(n - 1, 1))).to(bn.running_var.dtype), alpha=m)
        return y
    shape = (1, -1) + (1,) * (x.dim() - 2)
    scale = bn.weight * torch.rsqrt(bn.running_var + bn.eps)
    return x * scale.view(shape) + (bn.bias - bn.running_mean * scale).view(shape)


def warp_field(em_cl: torch.Tensor, theta: torch.Tensor, G: int = 64) -> torch.Tensor:
    """em_cl [N,E,E,E,3] fp32, theta [N,3,4] -> [N,3,G,G,G] (model.py:965-973)."""
    _chk_cuda(em_cl, torch.float32, "warp_field em")
    _chk_cuda(theta, torch.float32, "warp_field theta")
    N, E = em_cl.shape[0], em_cl.shape[1]
    out = torch.empty((N, 3, G, G, G), dtype=torch.float32, device=em_cl.device)
    L = _lib.load()
    _lib.check(L.mp_warp_field(_p(em_cl), _p(theta), _p(out), N, E, G, _stream()), "mp_warp_field")
    _count()
    return out


def warp_fused(v: Act, em_cl: torch.Tensor, theta: torch.Tensor, sum_d: bool, G: int = 64, f32: bool = True,
               split: bool = False) -> Act:
    """Fused flow->grid->trilinear gather on a CL volume (+ optional sum over D).  v.N may be 1 (shared source)."""
    _chk_cuda(em_cl, torch.float32, "warp_fused em")
    _chk_cuda(theta, torch.float32, "warp_fused theta")
    N, E = em_cl.shape[0], em_cl.shape[1]
    Nv, D, H, W, C = v.shape
    out = _alloc((N, 1 if sum_d else D, H, W, C), em_cl.device, f32, split)
    L = _lib.load()
    # algorithmic bytes: source volume(s) + 16^3 flow + the outputs actually written
    nbytes = Nv * D * H * W * C * 4 + N * E * E * E * 3 * 4
    nbytes += N * (1 if sum_d else D) * H * W * C * 4 * (int(f32) + int(split))
    with _Prof("warp_fused_sum" if sum_d else "warp_fused", 0, nbytes):
        _lib.check(L.mp_warp_fused_cl(_p(v.f32), _p(em_cl), _p(theta), _p(out.f32), _p(out.hi), _p(out.lo), N, Nv, C,
                                      D, H, W, E, G, 1 if sum_d else 0, _stream()), "mp_warp_fused_cl")
    _count()
    return out


@_profiled
def gn_relu_conv3x3_head(a: Act, ab: torch.Tensor, weight_host: torch.Tensor, bias_host: Optional[torch.Tensor],
                         act: int = ACT_SIGMOID) -> torch.Tensor:
    """act(conv3x3(relu(a * ab[...,0] + ab[...,1]))) in one pass: a CL fp32 [N,1,H,W,64], ab [N,64,2] (gn_finalize),
    weight_host (3,64,3,3) / bias_host (3,) contiguous fp32 CPU tensors -> NCHW fp32 [N,3,H,W] (model.py:748-751)."""
    N, D, H, W, C = a.shape
    if a.f32 is None or D != 1:
        raise RuntimeError("gn_relu_conv3x3_head: needs a 2-D fp32 channels-last activation")
    for t in (weight_host, bias_host):
        if t is not None and (t.is_cuda or t.dtype != torch.float32 or not t.is_contiguous()):
            raise RuntimeError("gn_relu_conv3x3_head: weights must be contiguous fp32 CPU tensors")
    Co = weight_host.shape[0]
    out = torch.empty((N, Co, H, W), dtype=torch.float32, device=a.device)
    L = _lib.load()
    _lib.check(L.mp_gn_relu_conv3x3_head(_p(a.f32), _p(ab), _p(weight_host), _p(bias_host), _p(out), N, H, W, C, Co, act,
                                         _stream()), "mp_gn_relu_conv3x3_head")
    _count()
    return out


def decode_jpeg_frames(blobs: Sequence[bytes], device="cuda") -> torch.Tensor:
    """JPEG bitstreams (host `bytes`, all the same size) -> uint8 RGB frames [N, H, W, 3] on `device` through nvJPEG
    (row f-4: `Image.open(...).convert("RGB")` of inference.py:10-13 for JPEG sources; only the compressed bytes cross PCIe)."""
    if not blobs:
        raise RuntimeError("decode_jpeg_frames: no frames")
    L = _lib.load()
    bufs = [ctypes.create_string_buffer(bytes(b), len(b)) for b in blobs]
    w, h = ctypes.c_int(), ctypes.c_int()
    _lib.check(L.mp_jpeg_info(ctypes.cast(bufs[0], ctypes.c_void_p), len(blobs[0]), ctypes.byref(w), ctypes.byref(h)), "mp_jpeg_info")
    n = len(blobs)
    out = torch.empty((n, h.value, w.value, 3), dtype=torch.uint8, device=device)
    ptrs = (ctypes.c_void_p * n)(*[ctypes.cast(b, ctypes.c_void_p) for b in bufs])
    sizes = (ctypes.c_size_t * n)(*[len(b) for b in blobs])
    with torch.cuda.device(out.device):
        _lib.check(L.mp_decode_jpeg_frames(ctypes.cast(ptrs, ctypes.c_void_p), ctypes.cast(sizes, ctypes.c_void_p), n, _p(out),
                                           h.value, w.value, _stream()), "mp_decode_jpeg_frames")
        torch.cuda.current_stream().synchronize()      # the host bitstream buffers are released when this function returns
    _count(n)
    return out


@_profiled
def frames_u8_to_f32(frames: torch.Tensor, mean: float = 0.5, std: float = 0.5) -> torch.Tensor:
    """uint8 HWC frames [N,H,W,3] on the device -> fp32 NCHW [N,3,H,W] = Normalize(mean, std)(ToTensor(frame))
    (reference inference.py:16-20)."""
    _chk_cuda(frames, torch.uint8, "frames_u8_to_f32")
    N, H, W, C = frames.shape
    if C != 3:
        raise RuntimeError("frames_u8_to_f32: frames must be [N, H, W, 3]")
    out = torch.empty((N, 3, H, W), dtype=torch.float32, device=frames.device)
    L = _lib.load()
    _lib.check(L.mp_frames_u8_to_f32(_p(frames), _p(out), N, H, W, mean, std, _stream()), "mp_frames_u8_to_f32")
    _count()
    return out


@_profiled
def frames_f32_to_u8(x: torch.Tensor, shift: float = 1.0, scale: float = 0.5, reverse_channels: bool = True) -> torch.Tensor:
    """fp32 NCHW [N,3,H,W] -> uint8 HWC [N,H,W,3] = ((x + shift) * scale * 255).astype(uint8), channels reversed like
    cv2.cvtColor(.., COLOR_BGR2RGB) (reference inference.py:36-43)."""
    _chk_cuda(x, torch.float32, "frames_f32_to_u8")
    N, C, H, W = x.shape
    if C != 3:
        raise RuntimeError("frames_f32_to_u8: x must be [N, 3, H, W]")
    out = torch.empty((N, H, W, 3), dtype=torch.uint8, device=x.device)
    L = _lib.load()
    _lib.check(L.mp_frames_f32_to_u8(_p(x), _p(out), N, H, W, shift, scale, 1 if reverse_channels else 0, _stream()),
               "mp_frames_f32_to_u8")
    _count()
    return out


@_profiled
def blur_subsample(x: torch.Tensor, kernel2d: torch.Tensor, step: int) -> torch.Tensor:
    _chk_cuda(x, torch.float32, "blur_subsample x")
    _chk_cuda(kernel2d, torch.float32, "blur_subsample kernel")
    N, C, H, W = x.shape
    ks = kernel2d.shape[-1]
    out = torch.empty((N, C, H // step, W // step), dtype=torch.float32, device=x.device)
    L = _lib.load()
    _lib.check(L.mp_blur_subsample(_p(x), _p(kernel2d), _p(out), N, C, H, W, ks, step, _stream()), "mp_blur_subsample")
    _count()
    return out
