"""libmpb200 inference plans for the motion encoder `Emtn` and the ResNet-50 descriptor branch
(row a7): the three conv trunks of `Emtn.forward` (model.py:888-907) -- two CIFAR-style ResNet-18s (resnet.py:160-310)
and the RepVGG-B1g2 deploy backbone of 6DRepNet (mysixdrepnet.py:30-69, 1215-1290) -- and `CustomResNet50`
(model.py:136-173) run on the same tcgen05 split-bf16 convolution kernels as the generator:

  * BatchNorm (eval) folded into the preceding convolution in float64;
  * RGB stems run on the tensor-core kernel with their 3 input channels zero-padded to 16;
  * stride-2 convolutions use TMA element strides, grouped convolutions one launch per group on channel windows;
  * MaxPool2d(3,2,1) and the global average pools are small HBM-bound kernels; the final Linear layers (<= 2048x512)
    are plain library GEMMs (`F.linear`).

Precision.  `Emtn`'s three trunks end in global average pools, which average the operand-rounding noise of the
convolutions away, so they run the two-pass fp16 convolution (`PREC_F16X2`: activations as ONE fp16 plane, weights as
fp16 hi + scaled fp16 lo, fp32 accumulation): against the CPU oracle R, t and z then differ by <= 2e-5 (relative L2;
tests/test_gpu_gbase.py holds them to 2e-4 of abs-max) while the convolutions need 2/3 of the tensor-pipe work and half
of the activation bytes.  Their 3x3 RGB stems run as a 1x1 convolution over 32 patch channels written by
`mp_im2col3x3_f16`.  `MPB200_EMTN_PREC=split` selects the three-pass split-bf16 plans instead (A/B runs); the
ResNet-50 descriptor (`es` feeds both warp generators un-pooled) always runs split-bf16.

Plans are built lazily from the live parameters, cached on parameter versions, and never registered as sub-modules
(the reference's state_dict keys are untouched).
"""
from __future__ import annotations

import os
from typing import List, Optional, Tuple

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import ops
from .ops import ACT_NONE, ACT_RELU, PREC_F16X2, PREC_SPLIT_BF16, Act

RGB_PAD = 16
EMTN_PREC = PREC_SPLIT_BF16 if os.environ.get("MPB200_EMTN_PREC", "f16x2") == "split" else PREC_F16X2
# fp16 plans: the CIFAR-ResNet stems (conv3x3 + BN + ReLU + max-pool) run as ONE kernel; 0 = im2col -> 1x1 conv -> max-pool (A/B)
STEM_FUSED = os.environ.get("MPB200_STEM_FUSED", "1") != "0"


def rgb16(x: torch.Tensor) -> Act:
    """NCHW fp32 [N,3,H,W] -> split channels-last activation with the channel axis zero-padded to 16."""
    return ops.from_nchw_pad16(x.contiguous())


def _bn(bn: nn.BatchNorm2d) -> dict:
    return {"weight": bn.weight.detach(), "bias": bn.bias.detach(), "running_mean": bn.running_mean,
            "running_var": bn.running_var}


def _pack_conv_bn(conv: nn.Conv2d, bn: Optional[nn.BatchNorm2d], cin_pad: int = 0,
                  prec: int = PREC_SPLIT_BF16) -> ops.PackedConv:
    w, b = conv.weight.detach(), (None if conv.bias is None else conv.bias.detach())
    if bn is not None:
        w, b = ops.fold_bn(w, b, _bn(bn), bn.eps)
    return ops.pack_conv(w, b, conv.weight.device, cin_pad=cin_pad, prec=prec)


class ResNetTrunkPlan:
    """torchvision-style ResNet trunk (BasicBlock or Bottleneck): stem conv+bn+relu, MaxPool(3,2,1), residual stages."""

    def __init__(self, conv1: nn.Conv2d, bn1: nn.BatchNorm2d, layers, prec: int = PREC_SPLIT_BF16):
        self.prec = prec
        self.stem_wb = ops.fold_bn(conv1.weight.detach(), None if conv1.bias is None else conv1.bias.detach(),
                                   _bn(bn1), bn1.eps)
        self.stem = None if prec == PREC_F16X2 else _pack_conv_bn(conv1, bn1, cin_pad=RGB_PAD)   # fp16: see DualResNet18Plan
        self.stem_stride = conv1.stride[0]
        self.blocks: List[Tuple] = []
        pk = lambda c, b: _pack_conv_bn(c, b, prec=prec)

        def pk_last(c, b, blk):
            """Last convolution of a block; a down-sampling shortcut (1x1 conv + BN, resnet.py:101-118) is folded and
            fused into its accumulator as extra K columns, so the shortcut tensor never exists in HBM."""
            if blk.downsample is None:
                return pk(c, b)
            w, bias = ops.fold_bn(c.weight.detach(), None if c.bias is None else c.bias.detach(), _bn(b), b.eps)
            dc, dbn = blk.downsample[0], blk.downsample[1]
            sc = ops.fold_bn(dc.weight.detach(), None if dc.bias is None else dc.bias.detach(), _bn(dbn), dbn.eps)
            return ops.pack_conv(w, bias, c.weight.device, prec=prec, shortcut=sc)

        for layer in layers:
            for blk in layer:
                ds = None if blk.downsample is None else blk.downsample[0].stride[0]     # stride of the fused shortcut
                if type(blk).__name__ == "BasicBlock":
                    convs = [(pk(blk.conv1, blk.bn1), blk.conv1.stride[0]), (pk_last(blk.conv2, blk.bn2, blk), 1)]
                else:   # Bottleneck (v1.5: the stride sits on conv2)
                    convs = [(pk(blk.conv1, blk.bn1), 1), (pk(blk.conv2, blk.bn2), blk.conv2.stride[0]),
                             (pk_last(blk.conv3, blk.bn3, blk), 1)]
                self.blocks.append((convs, ds))

    def features(self, x16: Act, last_f32: bool = False) -> Act:
        h, _ = ops.conv(x16, self.stem, act=ACT_RELU, f32=False, split=True, stride=self.stem_stride)
        h = ops.maxpool3x3s2(h)
        for bi, (convs, ds) in enumerate(self.blocks):
            t = h
            for pw, s in convs[:-1]:
                t, _ = ops.conv(t, pw, stride=s, act=ACT_RELU, f32=False, split=True)
            last = last_f32 and bi == len(self.blocks) - 1
            link = dict(res=h) if ds is None else dict(src2=h, stride2=ds)       # identity residual | fused 1x1 shortcut
            h, _ = ops.conv(t, convs[-1][0], stride=convs[-1][1], act=ACT_RELU, f32=last, split=not last, **link)
        return h

    def pooled(self, x16: Act) -> torch.Tensor:
        return ops.global_avgpool(self.features(x16))


class DualResNet18Plan:
    """The two CIFAR-style ResNet-18 trunks of Emtn (head pose, expression) read the same frame: their stems run as ONE
    3->128 convolution (+ one max-pool), and layer1 (64 channels each, no down-sampling) runs in lock-step on shared
    128-channel tensors through channel windows, so the 512x512 input and the 128-channel stem output are touched
    once instead of twice.  From layer2 on the two trunks are independent."""

    def __init__(self, a: ResNetTrunkPlan, b: ResNetTrunkPlan):
        assert a.prec == b.prec
        self.a, self.b, self.prec = a, b, a.prec
        (wa, ba), (wb, bb) = a.stem_wb, b.stem_wb
        self.c = wa.shape[0]
        w, bias = torch.cat([wa, wb], 0), torch.cat([ba, bb], 0)
        if self.prec == PREC_F16X2:
            self.stem = ops.pack_stem3x3_f16(w, bias, wa.device)      # 1x1 over the 32 im2col patch channels
        else:
            self.stem = ops.pack_conv(w, bias, wa.device, cin_pad=RGB_PAD)

    def pooled(self, x_in):
        """x_in: the NCHW fp32 frames (fp16 plans: stem + max-pool run as one kernel; with MPB200_STEM_FUSED=0 the fp16
        patch tensor of `ops.im2col3x3_f16(x, 1)`) or the padded split frame `rgb16(x)`."""
        c, half = self.c, self.prec == PREC_F16X2
        fmt = dict(f32=False, h16=True) if half else dict(f32=False, split=True)    # format of every intermediate
        if half and STEM_FUSED:
            h = ops.stem3x3_relu_maxpool_f16(x_in, self.stem)
        else:
            h, _ = ops.conv(x_in, self.stem, act=ACT_RELU, **fmt)
            h = ops.maxpool3x3s2_f16(h) if half else ops.maxpool3x3s2(h)
        n_lock = 0
        while (n_lock < len(self.a.blocks) and self.a.blocks[n_lock][1] is None and self.b.blocks[n_lock][1] is None
               and self.a.blocks[n_lock][0][0][1] == 1):
            out = ops._alloc(h.shape, h.device, False, not half, half)
            for j, pl in enumerate((self.a, self.b)):
                convs, _ds = pl.blocks[n_lock]
                t, _ = ops.conv(h, convs[0][0], act=ACT_RELU, in_c_off=j * c, **fmt)
                ops.conv(t, convs[1][0], res=h, act=ACT_RELU, out=out, out_c_off=j * c)
            h = out
            n_lock += 1
        feats = []
        for j, pl in enumerate((self.a, self.b)):
            g, off = h, j * c
            for bi in range(n_lock, len(pl.blocks)):
                convs, ds = pl.blocks[bi]
                win = off if bi == n_lock else 0
                t, _ = ops.conv(g, convs[0][0], stride=convs[0][1], act=ACT_RELU, in_c_off=win, **fmt)
                link = dict(res=g) if ds is None else dict(src2=g, stride2=ds, in2_c_off=win)
                g, _ = ops.conv(t, convs[1][0], act=ACT_RELU, **link, **fmt)
            feats.append(ops.global_avgpool_f16(g) if half else ops.global_avgpool(g))
        return feats


class RepVGGPlan:
    """RepVGG deploy backbone: 3x3 conv + ReLU blocks, stride 2 at stage heads, groups = 2 on even layers."""

    def __init__(self, backbone: nn.Module, prec: int = PREC_SPLIT_BF16):
        self.prec = prec
        self.blocks: List[Tuple] = []
        stages = [backbone.layer0] + [b for li in range(1, 5) for b in getattr(backbone, f"layer{li}")]
        for blk in stages:
            conv = blk.rbr_reparam
            g, s, cout = conv.groups, conv.stride[0], conv.out_channels
            w, b = conv.weight.detach(), conv.bias.detach()
            if g == 1 and conv.in_channels == 3 and prec == PREC_F16X2:
                packs = [ops.pack_stem3x3_f16(w, b, w.device)]       # layer0 on im2col patches: stride handled there
                s = 1
            elif g == 1:
                packs = [ops.pack_conv(w, b, w.device, cin_pad=RGB_PAD if conv.in_channels == 3 else 0, prec=prec)]
            else:
                cg = cout // g
                packs = [ops.pack_conv(w[i * cg:(i + 1) * cg], b[i * cg:(i + 1) * cg], w.device, prec=prec)
                         for i in range(g)]
            self.blocks.append((s, g, packs, cout))
        self.stem_stride = backbone.layer0.rbr_reparam.stride[0]

    def pooled(self, x_in) -> torch.Tensor:
        """x_in: `ops.im2col3x3_f16(x, self.stem_stride)` (fp16 plans) or `rgb16(x)`."""
        half = self.prec == PREC_F16X2
        fmt = dict(f32=False, h16=True) if half else dict(f32=False, split=True)
        h = x_in
        for s, g, packs, cout in self.blocks:
            if g == 1:
                h, _ = ops.conv(h, packs[0], stride=s, act=ACT_RELU, **fmt)
            else:
                N, D, H, W, C = h.shape
                out = ops._alloc((N, 1, H // s, W // s, cout), h.device, False, not half, half)
                for i, pw in enumerate(packs):
                    ops.conv(h, pw, stride=s, act=ACT_RELU, in_c_off=i * (C // g), out=out, out_c_off=i * (cout // g))
                h = out
        return ops.global_avgpool_f16(h) if half else ops.global_avgpool(h)


def _versions(*mods) -> tuple:
    s, dev = 0, None
    for m in mods:
        for t in list(m.parameters()) + list(m.buffers()):
            s += t._version + (t.data_ptr() & 0xFFFF)
            dev = t.device
    return (s, str(dev))


def emtn_plans(emtn):
    hp, ex, rot = emtn.head_pose_net, emtn.expression_net, emtn.rotation_net.model
    sig = _versions(hp, ex, rot) + (EMTN_PREC,)
    c = emtn.__dict__.get("_mp_cuda_plans")
    if c is None or c[0] != sig:
        with torch.no_grad():
            pa = ResNetTrunkPlan(hp.conv1, hp.bn1, [hp.layer1, hp.layer2, hp.layer3, hp.layer4], prec=EMTN_PREC)
            pb = ResNetTrunkPlan(ex[0], ex[1], [ex[4], ex[5], ex[6], ex[7]], prec=EMTN_PREC)
            c = (sig, (DualResNet18Plan(pa, pb), RepVGGPlan(rot, prec=EMTN_PREC)))
        emtn.__dict__["_mp_cuda_plans"] = c
    return c[1]


def emtn_forward(emtn, x: torch.Tensor):
    """Emtn.forward (model.py:888-907) on libmpb200 kernels: -> (Euler degrees [B,3], translation [B,3], z [B,512])."""
    from .emtn import ortho6d_to_euler_deg
    dual_plan, rot_plan = emtn_plans(emtn)
    x = x.float().contiguous()
    if dual_plan.prec == PREC_F16X2:
        in_rot = ops.im2col3x3_f16(x, rot_plan.stem_stride)
        in_dual = x if STEM_FUSED else ops.im2col3x3_f16(x, 1)
    else:
        in_rot = in_dual = rgb16(x)
    rot = emtn.rotation_net.model
    x6 = F.linear(rot_plan.pooled(in_rot), rot.linear_reg.weight, rot.linear_reg.bias)
    rotations = ortho6d_to_euler_deg(x6[:, :6])
    f_hp, f_ex = dual_plan.pooled(in_dual)
    head_pose = F.linear(f_hp, emtn.head_pose_net.fc.weight, emtn.head_pose_net.fc.bias)
    translation = head_pose[:, 3:]
    # AdaptiveAvgPool2d(1) then AdaptiveAvgPool2d((2,2)) (model.py:880-881): every channel replicated 2x2, NCHW flatten
    feat = f_ex.repeat_interleave(4, dim=1)
    expression = F.linear(feat, emtn.fc.weight, emtn.fc.bias)
    return rotations, translation, expression


def rotation_forward(emtn, x: torch.Tensor) -> torch.Tensor:
    """Only the 6DRepNet branch of `emtn_forward` (Euler degrees [B,3]) -- the differentiable / train-mode `Emtn` takes the angles
    from here and trains the two ResNet-18 trunks through the autograd Functions (emtn.py `_forward_autograd`)."""
    from .emtn import ortho6d_to_euler_deg
    rot = emtn.rotation_net.model
    sig = _versions(rot) + (EMTN_PREC,)
    c = emtn.__dict__.get("_mp_cuda_rot_plan")
    if c is None or c[0] != sig:
        with torch.no_grad():
            c = (sig, RepVGGPlan(rot, prec=EMTN_PREC))
        emtn.__dict__["_mp_cuda_rot_plan"] = c
    plan = c[1]
    x = x.float().contiguous()
    x_in = ops.im2col3x3_f16(x, plan.stem_stride) if plan.prec == PREC_F16X2 else rgb16(x)
    x6 = F.linear(plan.pooled(x_in), rot.linear_reg.weight, rot.linear_reg.bias)
    return ortho6d_to_euler_deg(x6[:, :6])


def resnet50_descriptor(r50, x: torch.Tensor) -> torch.Tensor:
    """CustomResNet50.forward (model.py:156-173) on libmpb200 kernels -> (B, 512, 2, 2) NCHW."""
    sig = _versions(r50)
    c = r50.__dict__.get("_mp_cuda_plan")
    if c is None or c[0] != sig:
        with torch.no_grad():
            c = (sig, (ResNetTrunkPlan(r50.conv1, r50.bn1, [r50.layer1, r50.layer2, r50.layer3]),
                       ops.pack_conv(r50.conv_reduce.weight, r50.conv_reduce.bias, r50.conv_reduce.weight.device)))
        r50.__dict__["_mp_cuda_plan"] = c
    trunk, reduce_pw = c[1]
    h = trunk.features(rgb16(x.float()), last_f32=True)           # [N,1,32,32,1024] fp32
    # AdaptiveAvgPool2d(2) on a 32x32 map == four successive 2x2 average pools (uniform blocks)
    while h.shape[2] > 2:
        h = ops.avgpool2(h, 1, f32=True, split=False)
    if h.shape[2] != 2 or h.shape[3] != 2:
        raise RuntimeError(f"CustomResNet50: unexpected feature map {h.shape} (expects 512x512 inputs)")
    out, _ = ops.conv(h, reduce_pw, f32=True)                      # 1x1 1024->512 on the 2x2 map
    return ops.to_nchw(out, 4)
