"""Host-side planning of the tensor-core convolution (tile boxes, swizzle widths, shared-memory / TMEM budgets) for every
convolution shape of the motion encoder in both operand formats -- `mp_conv_tc_supported` runs the planner only, so
this needs no GPU.  (A shape the planner rejects would surface on the GPU box as a RuntimeError, never as a fallback.)"""
import ctypes

import pytest

# (H, W, Cin, Cout, k, stride, in_C): Emtn's two CIFAR-style ResNet-18 trunks (resnet.py:160-310) ...
RESNET18 = [(512, 512, 32, 128, 1, 1, 32), (256, 256, 64, 64, 3, 1, 128), (256, 256, 64, 128, 3, 2, 128),
            (256, 256, 64, 128, 1, 2, 128), (128, 128, 128, 128, 3, 1, 128), (128, 128, 128, 256, 3, 2, 128),
            (128, 128, 128, 256, 1, 2, 128), (64, 64, 256, 256, 3, 1, 256), (64, 64, 256, 512, 3, 2, 256),
            (64, 64, 256, 512, 1, 2, 256), (32, 32, 512, 512, 3, 1, 512)]
# ... and RepVGG-B1g2 deploy (mysixdrepnet.py:1215-1290; groups = 2 on even layers -> half-width channel windows)
REPVGG = [(256, 256, 32, 64, 1, 1, 32), (256, 256, 64, 128, 3, 2, 64), (128, 128, 64, 64, 3, 1, 128),
          (128, 128, 128, 128, 3, 1, 128), (128, 128, 64, 128, 3, 2, 128), (64, 64, 256, 256, 3, 1, 256),
          (64, 64, 128, 128, 3, 1, 256), (64, 64, 256, 512, 3, 2, 256), (32, 32, 512, 512, 3, 1, 512),
          (32, 32, 256, 256, 3, 1, 512), (32, 32, 256, 1024, 3, 2, 512)]


@pytest.fixture(scope="module")
def L():
    from megaportrait_hack_b200 import lib
    lib.build()
    return lib


def _desc(lib, N, H, W, Cin, Cout, k, stride, in_C, prec):
    d = lib.ConvDesc()
    d.in_hi = d.in_lo = d.w_hi = d.w_lo = d.out_hi = d.out_lo = 256      # never dereferenced by the planner
    d.N, d.D, d.H, d.W, d.Cin, d.Cout = N, 1, H, W, Cin, Cout
    d.KD, d.KH, d.KW, d.Cout_pad = 1, k, k, (Cout + 15) // 16 * 16
    d.stride, d.in_C, d.out_C, d.prec = stride, in_C, Cout, prec
    return d


@pytest.mark.parametrize("N", [1, 32])
@pytest.mark.parametrize("prec", [0, 1], ids=["split_bf16", "f16x2"])
def test_motion_encoder_conv_shapes_are_planned(L, N, prec):
    so = L.load()
    for shape in RESNET18 + REPVGG:
        if prec == 0 and shape[2] == 32 and shape[4] == 1:
            continue        # the im2col stems exist in the fp16 plans only
        assert so.mp_conv_tc_supported(ctypes.byref(_desc(L, N, *shape, prec))) == 1, (N, shape, prec)


def test_planner_rejects_what_the_kernel_cannot_do(L):
    so = L.load()
    assert so.mp_conv_tc_supported(ctypes.byref(_desc(L, 1, 64, 64, 3, 64, 3, 1, 3, 0))) == 0      # Cin % 16
    assert so.mp_conv_tc_supported(ctypes.byref(_desc(L, 1, 64, 64, 64, 64, 3, 1, 64, 7))) == 0    # unknown prec
    assert so.mp_conv_tc_supported(ctypes.byref(_desc(L, 1, 63, 64, 64, 64, 3, 1, 64, 1))) == 0    # ragged grid


def _desc2(lib, N, H, W, Cin, Cout, k, prec, Cin2=0, stride2=1, in2_C=0, in2_off=0):
    d = _desc(lib, N, H, W, Cin, Cout, k, 1, Cin, prec)
    d.Cin2, d.stride2, d.in2_C, d.in2_c_off = Cin2, stride2, in2_C or Cin2, in2_off
    d.in2_hi = d.in2_lo = 256
    d.corr_scale = 2.0 ** -14
    return d


@pytest.mark.parametrize("N", [1, 32])
def test_fused_shortcut_and_fp8_shapes_are_planned(L, N):
    so = L.load()
    cases = [
        # G2d up-blocks (split-bf16): conv2 + fused 1x1 shortcut (model.py:616-640)
        (128, 128, 256, 256, 3, 0, 512, 1), (256, 256, 128, 128, 3, 0, 256, 1), (512, 512, 64, 64, 3, 0, 128, 1),
        # down-sampling ResNet-18 blocks of Emtn (fp16 two-pass), windowed / stride-2 second source
        (128, 128, 128, 128, 3, 1, 64, 2, 128, 64), (64, 64, 256, 256, 3, 1, 128, 2), (32, 32, 512, 512, 3, 1, 256, 2),
        # ResNet-50 bottlenecks of the descriptor branch: 1x1 conv3 + 1x1 shortcut
        (128, 128, 64, 256, 1, 0, 64, 1), (64, 64, 128, 512, 1, 0, 256, 2), (32, 32, 256, 1024, 1, 0, 512, 2),
        # G2d identity res-blocks and input conv in the fp16 + FP8 cross-term mode
        (64, 64, 512, 512, 3, 2), (64, 64, 64, 128, 1, 2),
    ]
    for c in cases:
        assert so.mp_conv_tc_supported(ctypes.byref(_desc2(L, N, *c))) == 1, (N, c)
    # the fp16 + FP8 format needs 64-channel groups; a shortcut must share the main operand's channel chunking
    assert so.mp_conv_tc_supported(ctypes.byref(_desc2(L, 1, 64, 64, 96, 128, 3, 2))) == 0
    assert so.mp_conv_tc_supported(ctypes.byref(_desc2(L, 1, 64, 64, 64, 64, 3, 0, 24, 1))) == 0


def test_operand_packing_accuracy():
    """Host packers of the three operand formats reproduce the weights to the advertised number of bits (CPU only)."""
    import torch
    from megaportrait_hack_b200 import ops
    g = torch.Generator().manual_seed(0)
    w = torch.randn(64, 128, 3, 3, generator=g) / 34
    wk = w.permute(0, 2, 3, 1).reshape(64, -1)
    p0 = ops.pack_conv(w, None)
    assert ((p0.w_hi.float() + p0.w_lo.float()) - wk).abs().max() <= wk.abs().max() * 2.0 ** -16
    # fp16 planes hold w * sw (sw = 1 / acc_scale, the power of two that puts the largest weight into [1, 2))
    p1 = ops.pack_conv(w, None, prec=ops.PREC_F16X2)
    assert 1.0 <= (wk / p1.acc_scale).abs().max() < 2.0
    assert ((p1.w_hi.float() + p1.w_lo.float() / ops.F16_LO_SCALE) * p1.acc_scale - wk).abs().max() <= wk.abs().max() * 2.0 ** -20
    p2 = ops.pack_conv(w, None, prec=ops.PREC_F16_Q8)
    hi = p2.w_hi.view(torch.float16).float()
    q = p2.w_lo.reshape(64, -1, 2, 64).view(torch.float8_e4m3fn).float()
    wl8, w8 = q[:, :, 0].reshape(64, -1), q[:, :, 1].reshape(64, -1)
    assert p2.w_hi.data_ptr() + p2.w_hi.numel() == p2.w_lo.data_ptr()          # one allocation: single-TMA [hi | lo] loads
    assert p2.corr_scale == 1.0 / ops.F16_LO_SCALE and 1.0 <= (wk / p2.acc_scale).abs().max() < 2.0
    assert ((hi + wl8 * p2.corr_scale) * p2.acc_scale - wk).abs().max() <= wk.abs().max() * 2.0 ** -15   # fp16 + e4m3 low part
    assert (w8 * p2.acc_scale - wk).abs().max() <= wk.abs().max() * 2.0 ** -4
    # tiny weights keep all their bits (no fp16 subnormals): the same relative accuracy at 2^-12 of the magnitude
    p3 = ops.pack_conv(w * 2.0 ** -12, None, prec=ops.PREC_F16X2)
    assert ((p3.w_hi.float() + p3.w_lo.float() / ops.F16_LO_SCALE) * p3.acc_scale - wk * 2.0 ** -12).abs().max() <= \
        wk.abs().max() * 2.0 ** -32
    # activation planes: fp16 plane + [e4m3(x) | e4m3((x - fp16 x) * 2048)] per 64-channel group
    x = torch.randn(2, 1, 3, 5, 128, generator=g) * 3
    qx = ops.q8_planes(x).reshape(2, 1, 3, 5, 2, 2, 64).view(torch.float8_e4m3fn).float()
    x8, xl8 = qx[..., 0, :].reshape(x.shape), qx[..., 1, :].reshape(x.shape)
    assert (x8 - x).abs().max() <= x.abs().max() * 2.0 ** -4
    assert ((x.half().float() + xl8 / 2048) - x).abs().max() <= x.abs().max() * 2.0 ** -15
    # per-tensor power-of-two scale of the byte plane: large activations no longer saturate e4m3 (448)
    xb = x * 300.0
    sc = ops.q8_scale_for(xb.abs().max().item())
    qs = ops.q8_planes(xb, sc).reshape(2, 1, 3, 5, 2, 2, 64).view(torch.float8_e4m3fn).float()
    assert (qs[..., 0, :].reshape(x.shape) / sc - xb).abs().max() <= xb.abs().max() * 2.0 ** -4
    q1 = ops.q8_planes(xb, 1.0).reshape(2, 1, 3, 5, 2, 2, 64).view(torch.float8_e4m3fn).float()
    assert (q1[..., 0, :].reshape(x.shape) - xb).abs().max() > xb.abs().max() * 0.25      # unscaled: clipped at 448
