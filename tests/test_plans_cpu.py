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
