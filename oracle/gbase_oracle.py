"""TEST INFRASTRUCTURE ONLY -- CPU fp32 restatement of the reference's Gbase forward path.

This file is the parity checker for the CUDA path.  Only `tests/`, `__graft_entry__.smoke()` and
`bench.py`'s CPU-baseline / `--impl reference` legs may import it; the product package never does.

It restates `/root/reference/model.py` (commit 580cab3) as pure functions over a flat `state_dict`
(`key -> tensor`, the reference's own 971 keys + the un-registered 6DRepNet tensors under
`motionEncoder.rotation_net.model.*`).  Arithmetic is delegated to the same third-party library the reference
calls, PyTorch ATen on CPU (torch 2.11.0 in this image; the reference pins no version, requirements.txt:1-20):
`F.conv2d/3d`, `F.group_norm`, `F.batch_norm`, `F.interpolate`, `F.affine_grid`, `F.grid_sample`.

Parity status: PINNED against the real reference run in the authoring container -- `oracle/make_golden.py`
imports `/root/reference/model.py` through `oracle/ref_shim.py`, fills it with the per-key seeded weights
(`megaportrait_hack_b200/seeded.py`) and stores per-stage outputs in `tests/golden/`;
`tests/test_oracle_golden.py` checks this restatement against them (the reference itself ships no tests,
golden vectors or checkpoints: SURVEY.md section 4 / 8c).

Each function cites the reference lines it follows.  Eval-mode semantics (BatchNorm uses running stats) unless `BN_TRAINING` is set.
"""
from __future__ import annotations

import math
from typing import Dict, Tuple

import torch
import torch.nn.functional as F

SD = Dict[str, torch.Tensor]
ROTNET_PREFIX = "motionEncoder.rotation_net.model."


# ----------------------------------------------------------------------------- small helpers
def _conv(x, sd: SD, p: str, stride=1, padding=0, groups=1):
    w = sd[p + ".weight"]
    b = sd.get(p + ".bias")
    fn = F.conv3d if w.dim() == 5 else F.conv2d
    return fn(x, w, b, stride=stride, padding=padding, groups=groups)


BN_TRAINING = False     # True: nn.BatchNorm2d as in `.train()` (batch statistics) -- the gradient tests of SURVEY.md row f-2


def _bn(x, sd: SD, p: str):
    """nn.BatchNorm2d, eps 1e-5 (model.py:605-616 and the resnets): running statistics (eval mode, the default) or, under
    `BN_TRAINING`, the batch statistics `Gbase.train()` uses (train.py:133)."""
    if BN_TRAINING:
        return F.batch_norm(x, None, None, sd[p + ".weight"], sd[p + ".bias"], training=True, eps=1e-5)
    return F.batch_norm(x, sd[p + ".running_mean"], sd[p + ".running_var"], sd[p + ".weight"], sd[p + ".bias"],
                        training=False, eps=1e-5)


def _gn(x, sd: SD, p: str, groups=32):
    return F.group_norm(x, groups, sd[p + ".weight"], sd[p + ".bias"], eps=1e-5)


def conv2d_ws(x, sd: SD, p: str, padding=1):
    """Conv2d_WS.forward, model.py:61-69: per-out-channel (w-mean)/(unbiased std + 1e-5)."""
    w = sd[p + ".weight"]
    mean = w.mean(dim=(1, 2, 3), keepdim=True)
    w = w - mean
    std = w.reshape(w.size(0), -1).std(dim=1).view(-1, 1, 1, 1) + 1e-5
    w = w / std
    return F.conv2d(x, w, sd[p + ".bias"], padding=padding)


def resblock_custom2d(x, sd: SD, p: str):
    """ResBlock_Custom.forward (dimension=2), model.py:110-130."""
    out2 = _conv(x, sd, p + ".conv_res", padding=1)
    out1 = F.relu(F.group_norm(x, 32))
    out1 = conv2d_ws(out1, sd, p + ".conv_ws")
    out1 = F.relu(F.group_norm(out1, 32))
    out1 = _conv(out1, sd, p + ".conv", padding=1)
    return out1 + out2


def adaptive_group_norm(x, sd: SD, p: str):
    """AdaptiveGroupNorm.forward, model.py:314-316."""
    return _gn(x, sd, p + ".group_norm") * sd[p + ".weight"] + sd[p + ".bias"]


def resblock3d_adaptive(x, sd: SD, p: str):
    """ResBlock3D_Adaptive.forward, model.py:385-408 (upsample is False at every call site)."""
    out = _conv(x, sd, p + ".conv1", padding=1)
    out = F.relu(adaptive_group_norm(out, sd, p + ".norm1"))
    out = _conv(out, sd, p + ".conv2", padding=1)
    out = adaptive_group_norm(out, sd, p + ".norm2")
    res = _conv(x, sd, p + ".residual_conv") if (p + ".residual_conv.weight") in sd else x
    return F.relu(out + res)


def _bottleneck(x, sd: SD, p: str, stride: int):
    """torchvision Bottleneck (v1.5: stride on conv2), used by CustomResNet50, model.py:139-146."""
    out = F.relu(_bn(_conv(x, sd, p + ".conv1"), sd, p + ".bn1"))
    out = F.relu(_bn(_conv(out, sd, p + ".conv2", stride=stride, padding=1), sd, p + ".bn2"))
    out = _bn(_conv(out, sd, p + ".conv3"), sd, p + ".bn3")
    if (p + ".downsample.0.weight") in sd:
        x = _bn(_conv(x, sd, p + ".downsample.0", stride=stride), sd, p + ".downsample.1")
    return F.relu(out + x)


def custom_resnet50(x, sd: SD, p: str):
    """CustomResNet50.forward, model.py:156-173."""
    x = F.relu(_bn(_conv(x, sd, p + ".conv1", stride=2, padding=3), sd, p + ".bn1"))
    x = F.max_pool2d(x, 3, 2, 1)
    for li, (nblk, stride) in enumerate(((3, 1), (4, 2), (6, 2)), start=1):
        for b in range(nblk):
            x = _bottleneck(x, sd, f"{p}.layer{li}.{b}", stride if b == 0 else 1)
    x = F.adaptive_avg_pool2d(x, 2)
    return _conv(x, sd, p + ".conv_reduce")


EAPP_3D_ORDER = ("resblock3D_96", "resblock3D_96_2", "resblock3D_96_1", "resblock3D_96_1_2",
                 "resblock3D_96_2", "resblock3D_96_2_2")  # model.py:276-290 (resblock3D_96_2 runs twice)


def eapp(x, sd: SD, p: str = "appearanceEncoder") -> Tuple[torch.Tensor, torch.Tensor]:
    """Eapp.forward, model.py:245-299."""
    out = _conv(x, sd, p + ".conv", padding=3)
    for name in ("resblock_128", "resblock_256", "resblock_512"):
        out = resblock_custom2d(out, sd, f"{p}.{name}")
        out = F.avg_pool2d(out, 2, 2)
    out = F.relu(F.group_norm(out, 32))
    out = _conv(out, sd, p + ".conv_1")
    vs = out.view(out.size(0), 96, 16, *out.shape[2:])
    for name in EAPP_3D_ORDER:
        vs = resblock3d_adaptive(vs, sd, f"{p}.{name}")
    es = custom_resnet50(x, sd, p + ".custom_resnet50")
    es = F.linear(torch.flatten(es, 1), sd[p + ".fc.weight"], sd[p + ".fc.bias"])
    return vs, es


# ----------------------------------------------------------------------------- Emtn (model.py:869-907)
def _basic_block(x, sd: SD, p: str, stride: int):
    out = F.relu(_bn(_conv(x, sd, p + ".conv1", stride=stride, padding=1), sd, p + ".bn1"))
    out = _bn(_conv(out, sd, p + ".conv2", padding=1), sd, p + ".bn2")
    if (p + ".downsample.0.weight") in sd:
        x = _bn(_conv(x, sd, p + ".downsample.0", stride=stride), sd, p + ".downsample.1")
    return F.relu(out + x)


def _resnet18_trunk(x, sd: SD, p: str, names):
    """CIFAR-style ResNet-18 trunk (resnet.py:192-204,282-291): 3x3 stride-1 stem, maxpool, 4 stages, GAP."""
    conv1, bn1, layers = names
    x = F.relu(_bn(_conv(x, sd, f"{p}.{conv1}", padding=1), sd, f"{p}.{bn1}"))
    x = F.max_pool2d(x, 3, 2, 1)
    for li, lname in enumerate(layers):
        for b in range(2):
            x = _basic_block(x, sd, f"{p}.{lname}.{b}", 2 if (li > 0 and b == 0) else 1)
    return F.adaptive_avg_pool2d(x, 1)


def ortho6d_to_euler_deg(x6: torch.Tensor) -> torch.Tensor:
    """mysixdrepnet.py:272-315 then *180/pi (mysixdrepnet.py:826-828)."""
    def _norm(v):
        mag = torch.sqrt(v.pow(2).sum(1)).clamp_min(1e-8)
        return v / mag[:, None]
    xr, yr = x6[:, 0:3], x6[:, 3:6]
    x = _norm(xr)
    z = _norm(torch.cross(x, yr, dim=1))
    y = torch.cross(z, x, dim=1)
    R = torch.stack((x, y, z), dim=2)
    sy = torch.sqrt(R[:, 0, 0] ** 2 + R[:, 1, 0] ** 2)
    singular = (sy < 1e-6).float()
    ex = torch.atan2(R[:, 2, 1], R[:, 2, 2])
    ey = torch.atan2(-R[:, 2, 0], sy)
    ez = torch.atan2(R[:, 1, 0], R[:, 0, 0])
    exs = torch.atan2(-R[:, 1, 2], R[:, 1, 1])
    eys = torch.atan2(-R[:, 2, 0], sy)
    ezs = R[:, 1, 0] * 0
    e = torch.stack((ex * (1 - singular) + exs * singular, ey * (1 - singular) + eys * singular,
                     ez * (1 - singular) + ezs * singular), dim=1)
    return e * 180 / math.pi


def sixdrepnet_euler(x, sd: SD, p: str = ROTNET_PREFIX[:-1]) -> torch.Tensor:
    """MySixDRepNet('RepVGG-B1g2', deploy=True).forward + SixDRepNet_Detector.predict
    (mysixdrepnet.py:30-69, 802-833, 1085-1120, 1215-1290): plain 3x3-conv+ReLU stages, stride 2 at the head of
    each, groups=2 on even layer indices, GAP, Linear(2048,6), 6-D -> rotation matrix -> Euler degrees."""
    def block(x, q, stride):
        w = sd[q + ".rbr_reparam.weight"]
        groups = x.shape[1] // w.shape[1]
        return F.relu(F.conv2d(x, w, sd[q + ".rbr_reparam.bias"], stride=stride, padding=1, groups=groups))
    x = block(x, p + ".layer0", 2)
    for li, nblk in enumerate((4, 6, 16, 1), start=1):
        for b in range(nblk):
            x = block(x, f"{p}.layer{li}.{b}", 2 if b == 0 else 1)
    x = torch.flatten(F.adaptive_avg_pool2d(x, 1), 1)
    x = F.linear(x, sd[p + ".linear_reg.weight"], sd[p + ".linear_reg.bias"])
    return ortho6d_to_euler_deg(x[:, :6])


def emtn(x, sd: SD, p: str = "motionEncoder", rot=None):
    """Emtn.forward, model.py:888-907: (Euler degrees from 6DRepNet, translation = head_pose[:,3:], z).  `rot` (gradient tests):
    angles to use instead of running 6DRepNet."""
    if rot is None:
        rot = sixdrepnet_euler(x, sd)
    hp = _resnet18_trunk(x, sd, p + ".head_pose_net", ("conv1", "bn1", ("layer1", "layer2", "layer3", "layer4")))
    hp = F.linear(torch.flatten(hp, 1), sd[p + ".head_pose_net.fc.weight"], sd[p + ".head_pose_net.fc.bias"])
    t = hp[:, 3:]
    ex = _resnet18_trunk(x, sd, p + ".expression_net", ("0", "1", ("4", "5", "6", "7")))
    ex = F.adaptive_avg_pool2d(ex, (2, 2))  # appended after the GAP, model.py:880-881
    z = F.linear(torch.flatten(ex, 1), sd[p + ".fc.weight"], sd[p + ".fc.bias"])
    return rot, t, z


# ----------------------------------------------------------------------------- warp generators
def flowfield(zs, sd: SD, p: str):
    """FlowField.forward, model.py:439-471 (adaptive_gamma/beta are ignored by the reference)."""
    x = _conv(zs, sd, p + ".conv1x1")
    x = x.view(-1, 512, 4, *x.shape[2:])
    for i, sf in enumerate(((2, 2, 2), (2, 2, 2), (1, 2, 2), (1, 2, 2)), start=1):
        x = resblock3d_adaptive(x, sd, f"{p}.resblock{i}")
        x = F.interpolate(x, scale_factor=sf, mode="nearest")
    x = _conv(x, sd, p + ".conv3x3x3", padding=1)
    x = F.group_norm(x, 1, sd[p + ".gn.weight"], sd[p + ".gn.bias"], eps=1e-5)
    return torch.tanh(F.relu(x))


def rotation_matrix(rot_deg):
    """compute_rotation_matrix, model.py:811-856: R = Rx(a) @ (Ry(b) @ Rz(c)), degrees in."""
    r = rot_deg * (math.pi / 180.0)
    ca, sa, cb, sb, cg, sg = torch.cos(r[:, 0]), torch.sin(r[:, 0]), torch.cos(r[:, 1]), torch.sin(r[:, 1]), \
        torch.cos(r[:, 2]), torch.sin(r[:, 2])
    z, o = torch.zeros_like(ca), torch.ones_like(ca)
    Rx = torch.stack((torch.stack((o, z, z), 1), torch.stack((z, ca, -sa), 1), torch.stack((z, sa, ca), 1)), 1)
    Ry = torch.stack((torch.stack((cb, z, sb), 1), torch.stack((z, o, z), 1), torch.stack((-sb, z, cb), 1)), 1)
    Rz = torch.stack((torch.stack((cg, -sg, z), 1), torch.stack((sg, cg, z), 1), torch.stack((z, z, o), 1)), 1)
    return Rx @ (Ry @ Rz)


def affine_3x4(rot_deg, trans, invert: bool):
    """First three rows of the 4x4 [R|t] (optionally inverted), model.py:790-803."""
    B = rot_deg.shape[0]
    A = torch.eye(4).repeat(B, 1, 1)
    A[:, :3, :3] = rotation_matrix(rot_deg)
    A[:, :3, 3] = trans
    if invert:
        A = torch.inverse(A)
    return A[:, :3]


def rt_warp(rot_deg, trans, invert: bool, grid_size=64):
    """compute_rt_warp, model.py:777-809."""
    theta = affine_3x4(rot_deg, trans, invert)
    grid = F.affine_grid(theta, (rot_deg.shape[0], 1, grid_size, grid_size, grid_size), align_corners=False)
    return grid.permute(0, 4, 1, 2, 3)


def warp_generator(R, t, z, e, sd: SD, p: str, invert: bool):
    """WarpGeneratorS2C.forward (invert=True) / WarpGeneratorC2D.forward (invert=False), model.py:938-1024."""
    assert R.shape == (z.shape[0], 3) and t.shape == (z.shape[0], 3) and z.shape == e.shape
    s = torch.matmul(z + e, sd[p + ".adaptive_matrix_gamma"])
    w_em = flowfield(s[:, :, None, None], sd, p + ".flowfield")
    w_rt = rt_warp(R, t, invert, 64)
    w_em = F.interpolate(w_em, size=w_rt.shape[2:], mode="trilinear", align_corners=False)
    return w_rt + w_em, w_em


def apply_warping_field(v, warp_field):
    """apply_warping_field, model.py:1028-1065 (including its re-normalisation, :1056-1058)."""
    B, C, D, H, W = v.shape
    warp_field = F.interpolate(warp_field, size=(D, H, W), mode="trilinear", align_corners=True)
    d = torch.linspace(-1, 1, D)
    h = torch.linspace(-1, 1, H)
    w = torch.linspace(-1, 1, W)
    gd, gh, gw = torch.meshgrid(d, h, w, indexing="ij")
    grid = torch.stack((gw, gh, gd), dim=-1)[None].repeat(B, 1, 1, 1, 1)
    g = grid + warp_field.permute(0, 2, 3, 4, 1)
    g = 2.0 * g / torch.tensor([W - 1, H - 1, D - 1]) - 1.0
    return F.grid_sample(v, g, mode="bilinear", padding_mode="border", align_corners=True)


# ----------------------------------------------------------------------------- G3d / G2d / pyramid
def resblock3d(x, sd: SD, p: str):
    """ResBlock3D.forward, model.py:512-528."""
    idt = _conv(x, sd, p + ".shortcut") if (p + ".shortcut.weight") in sd else x
    out = F.relu(_gn(_conv(x, sd, p + ".conv1", padding=1), sd, p + ".gn1"))
    out = _gn(_conv(out, sd, p + ".conv2", padding=1), sd, p + ".gn2")
    return F.relu(out + idt)


def g3d(x, sd: SD, p: str = "G3d"):
    """G3d.forward, model.py:593-597 (Sequential indices 0,2,4,6 / 0,2,4 hold the blocks, :574-590)."""
    for i in (0, 2, 4, 6):
        x = resblock3d(x, sd, f"{p}.downsampling.{i}")
        if i < 6:
            x = F.avg_pool3d(x, 2, 2)
    for i in (0, 2, 4):
        x = resblock3d(x, sd, f"{p}.upsampling.{i}")
        x = F.interpolate(x, scale_factor=2, mode="trilinear", align_corners=True)
    return _conv(x, sd, p + ".final_conv", padding=1)


def resblock2d(x, sd: SD, p: str):
    """ResBlock2D.forward, model.py:621-640 (downsample is False at every call site)."""
    out = F.relu(_bn(_conv(x, sd, p + ".conv1", padding=1), sd, p + ".bn1"))
    out = _bn(_conv(out, sd, p + ".conv2", padding=1), sd, p + ".bn2")
    if (p + ".shortcut.0.weight") in sd:
        x = _bn(_conv(x, sd, p + ".shortcut.0"), sd, p + ".shortcut.1")
    return F.relu(out + x)


def g2d(x, sd: SD, p: str = "G2d"):
    """G2d.forward, model.py:754-763."""
    x = _conv(x, sd, p + ".reshape")
    x = _conv(x, sd, p + ".conv1x1")
    for i in range(8):
        x = resblock2d(x, sd, f"{p}.res_blocks.{i}")
    for name in ("upsample1", "upsample2", "upsample3"):
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
        x = resblock2d(x, sd, f"{p}.{name}.1")
    x = F.relu(_gn(x, sd, p + ".final_conv.0"))
    return torch.sigmoid(_conv(x, sd, p + ".final_conv.2", padding=1))


def gaussian_kernel(scale: float, channels: int = 3):
    """AntiAliasInterpolation2d.__init__, model.py:650-681."""
    sigma = (1 / scale - 1) / 2
    ks = 2 * round(sigma * 4) + 1
    ax = torch.arange(ks, dtype=torch.float32)
    mean = (ks - 1) / 2
    g1 = torch.exp(-(ax - mean) ** 2 / (2 * sigma ** 2))
    k = g1[:, None] * g1[None, :]
    k = k / k.sum()
    return k.view(1, 1, ks, ks).repeat(channels, 1, 1, 1), ks // 2


def image_pyramid(x, scales=(0.5, 0.25)):
    """ImagePyramide.forward / AntiAliasInterpolation2d.forward, model.py:683-691, 1081-1085."""
    out = {}
    for s in scales:
        k, ka = gaussian_kernel(s, x.shape[1])
        y = F.conv2d(F.pad(x, (ka, ka, ka, ka)), k, groups=x.shape[1])
        out["prediction_" + str(s)] = F.interpolate(y, scale_factor=(s, s))
    return out


# ----------------------------------------------------------------------------- Gbase
def encode_source(xs, sd: SD):
    """Source-only half of Gbase.forward, model.py:1141-1160."""
    st = {}
    st["vs"], st["es"] = eapp(xs, sd)
    st["Rs"], st["ts"], st["zs"] = emtn(xs, sd)
    st["w_s2c"], st["w_em_s2c"] = warp_generator(st["Rs"], st["ts"], st["zs"], st["es"], sd,
                                                 "warp_generator_s2c", invert=True)
    st["vc"] = apply_warping_field(st["vs"], st["w_s2c"])
    assert st["vc"].shape[1:] == (96, 16, 64, 64)
    st["vc2d"] = g3d(st["vc"], sd)
    return st


def drive(src: dict, xd, sd: SD):
    """Driver half of Gbase.forward, model.py:1145, 1163-1180."""
    st = {}
    st["Rd"], st["td"], st["zd"] = emtn(xd, sd)
    st["w_c2d"], st["w_em_c2d"] = warp_generator(st["Rd"], st["td"], st["zd"], src["es"], sd,
                                                 "warp_generator_c2d", invert=False)
    st["warped"] = apply_warping_field(src["vc2d"], st["w_c2d"])
    assert st["warped"].shape[1:] == (96, 16, 64, 64)
    st["projected"] = torch.sum(st["warped"], dim=2)
    st["rgb"] = g2d(st["projected"], sd)
    st["pyramids"] = image_pyramid(st["rgb"])
    return st


@torch.no_grad()
def gbase_forward(xs, xd, sd: SD, stages: bool = False):
    """Gbase.forward(xs, xd) -> (xhat_base, pyramids), model.py:1140-1180.  Bs must equal Bd (model.py:993)."""
    src = encode_source(xs, sd)
    drv = drive(src, xd, sd)
    if stages:
        return drv["rgb"], drv["pyramids"], {**src, **drv}
    return drv["rgb"], drv["pyramids"]


def gbase_forward_train(xs, xd, sd: SD, rotations=None):
    """`Gbase.forward` in the reference's statement order (model.py:1141-1178) WITH autograd recording and, under `BN_TRAINING`,
    train-mode BatchNorm -- the checker of the differentiable path (train.py:194).  -> (rgb, pyramids, stages).
    `rotations` = (Rs, Rd) replaces 6DRepNet's output (its weights are outside `Gbase.parameters()` and never trained)."""
    st = {}
    st["vs"], st["es"] = eapp(xs, sd)
    rs, rd = rotations if rotations is not None else (None, None)
    st["Rs"], st["ts"], st["zs"] = emtn(xs, sd, rot=rs)
    st["Rd"], st["td"], st["zd"] = emtn(xd, sd, rot=rd)
    st["w_s2c"], _ = warp_generator(st["Rs"], st["ts"], st["zs"], st["es"], sd, "warp_generator_s2c", invert=True)
    st["vc"] = apply_warping_field(st["vs"], st["w_s2c"])
    st["vc2d"] = g3d(st["vc"], sd)
    st["w_c2d"], _ = warp_generator(st["Rd"], st["td"], st["zd"], st["es"], sd, "warp_generator_c2d", invert=False)
    st["warped"] = apply_warping_field(st["vc2d"], st["w_c2d"])
    st["projected"] = torch.sum(st["warped"], dim=2)
    st["rgb"] = g2d(st["projected"], sd)
    st["pyramids"] = image_pyramid(st["rgb"])
    return st["rgb"], st["pyramids"], st


@torch.no_grad()
def gbase_forward_shared_source(xs1, xd, sd: SD, stages: bool = False):
    """`Gbase(xs.expand(N), xd)` computed with the source half evaluated once (eval mode has no cross-sample
    coupling): the semantics of BASELINE configs 2-3."""
    src = encode_source(xs1, sd)
    n = xd.shape[0]
    srcN = {k: (v.expand(n, *v.shape[1:]) if torch.is_tensor(v) else v) for k, v in src.items()}
    drv = drive(srcN, xd, sd)
    if stages:
        return drv["rgb"], drv["pyramids"], {**src, **drv}
    return drv["rgb"], drv["pyramids"]


# ----------------------------------------------------------------------------- Genh / GHR (row f-3)
def genh(x, sd: SD, p: str = "Genh"):
    """Genh.forward, model.py:1349-1391, with `ResBlock2D(64)` read as `ResBlock2D(64, 64)` (the reference raises TypeError as
    written, model.py:1354 vs 601).  Sequential indices: encoder 0 conv7x7, 1/3/5/7 blocks, 2/4/6 AvgPool2d(2); res_blocks
    0..7; decoder 0/2/4 bilinear x2 (align_corners=True), 1/3/5 blocks, 6 conv7x7, 7 tanh."""
    pre = (p + ".") if p else ""
    x = _conv(x, sd, pre + "encoder.0", padding=3)
    for i in (1, 3, 5, 7):
        x = resblock2d(x, sd, f"{pre}encoder.{i}")
        if i < 7:
            x = F.avg_pool2d(x, 2, 2)
    for i in range(8):
        x = resblock2d(x, sd, f"{pre}res_blocks.{i}")
    for i in (1, 3, 5):
        x = F.interpolate(x, scale_factor=2, mode="bilinear", align_corners=True)
        x = resblock2d(x, sd, f"{pre}decoder.{i}")
    return torch.tanh(_conv(x, sd, pre + "decoder.6", padding=3))


def ghr_forward(xs, xd, sd_gbase: SD, sd_genh: SD):
    """GHR.forward, model.py:1446-1450, with the tuple returned by Gbase.forward reduced to its image (model.py:1180)."""
    rgb, _ = gbase_forward(xs, xd, sd_gbase)
    return genh(rgb, sd_genh, "")
