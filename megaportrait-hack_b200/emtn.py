"""Motion encoder `Emtn` (model.py:869-907) and the `CustomResNet50` descriptor branch
(model.py:136-173, row a7): module definitions with the reference's attribute names and state_dict keys.

`backend = "mpb200"` (the default) runs both on libmpb200 kernels (`emtn_cuda.py`): CUDA tensors, inference only --
CPU tensors, train mode and autograd RAISE (no silent ATen fallback).  `backend = "cudnn"` is an explicit opt-in to the
stock ATen / cuDNN graph (BatchNorm folded, channels_last; plain module graph in train mode or under autograd): the GPU
incumbent for A/B timing and the second opinion of tests/test_gpu_gbase.py; it is never selected automatically.
"""
from __future__ import annotations

import math

import torch
import torch.nn as nn
import torch.nn.functional as F
import torchvision.models as tvm
from torchvision.models.resnet import BasicBlock, ResNet

FEATURE_SIZE_AVG_POOL = 2   # model.py:46
FEATURE_SIZE = (2, 2)       # model.py:47
COMPRESS_DIM = 512          # model.py:48


def cifar_resnet18(num_classes: int = 10) -> ResNet:
    """resnet.py:160-310 `resnet18`: torchvision BasicBlock ResNet-18 with a 3x3 stride-1 stem (resnet.py:192)."""
    m = ResNet(BasicBlock, [2, 2, 2, 2], num_classes=num_classes)
    m.conv1 = nn.Conv2d(3, 64, kernel_size=3, stride=1, padding=1, bias=False)
    nn.init.kaiming_normal_(m.conv1.weight, mode="fan_out", nonlinearity="relu")
    return m


class CustomResNet50(nn.Module):
    """model.py:136-173: torchvision resnet50 trunk up to layer3 -> AdaptiveAvgPool(2) -> 1x1 conv 1024->512."""

    def __init__(self, *args, **kwargs):
        super().__init__()
        resnet = tvm.resnet50(*args, **kwargs)
        self.conv1 = resnet.conv1
        self.bn1 = resnet.bn1
        self.maxpool = resnet.maxpool
        self.layer1 = resnet.layer1
        self.layer2 = resnet.layer2
        self.layer3 = resnet.layer3
        self.adaptive_avg_pool = nn.AdaptiveAvgPool2d(FEATURE_SIZE_AVG_POOL)
        self.conv_reduce = nn.Conv2d(1024, 512, kernel_size=1)
        self.backend = "mpb200"

    def _forward_autograd(self, x):
        """Differentiable / train-mode form: every convolution and BatchNorm on the libmpb200 autograd Functions."""
        x = resnet_trunk_autograd(self.conv1, self.bn1, self.maxpool, (self.layer1, self.layer2, self.layer3), x)
        from . import ops
        return ops.conv_train(self.adaptive_avg_pool(x), self.conv_reduce.weight, self.conv_reduce.bias)

    def forward(self, x):
        if getattr(self, "backend", "mpb200") == "mpb200":
            if x.is_cuda and (self.training or _wants_grad(self, x)):
                return self._forward_autograd(x.float()).contiguous()
            _require_mpb200(self, x)
            from . import emtn_cuda
            return emtn_cuda.resnet50_descriptor(self, x)
        if self.training or torch.is_grad_enabled():        # backend == "cudnn" (explicit opt-in): stock module graph
            x = F.relu(self.bn1(self.conv1(x)))
            x = self.maxpool(x)
            x = self.layer3(self.layer2(self.layer1(x)))
        else:
            sig = _versions(self)
            c = self.__dict__.get("_mp_plan")
            if c is None or c[0] != sig:
                with torch.no_grad():
                    c = (sig, _FoldedResNetTrunk(self.conv1, self.bn1, self.maxpool,
                                                 [self.layer1, self.layer2, self.layer3]))
                self.__dict__["_mp_plan"] = c
            x = c[1](x)
        x = self.adaptive_avg_pool(x)
        return self.conv_reduce(x)


class _RepVGGDeployBlock(nn.Module):
    """RepVGGBlock in deploy form (mysixdrepnet.py:1085-1120): one re-parameterised 3x3 conv + ReLU."""

    def __init__(self, cin, cout, stride, groups):
        super().__init__()
        self.rbr_reparam = nn.Conv2d(cin, cout, 3, stride=stride, padding=1, groups=groups, bias=True)

    def forward(self, x):
        return F.relu(self.rbr_reparam(x))


class SixDRepNetBackbone(nn.Module):
    """MySixDRepNet('RepVGG-B1g2', deploy=True) (mysixdrepnet.py:30-69, 1215-1290): stages of 1/4/6/16/1 blocks,
    widths 64/128/256/512/2048, stride 2 at the head of each stage, groups=2 on even layer indices (g2_map)."""

    def __init__(self):
        super().__init__()
        widths, blocks = (128, 256, 512, 2048), (4, 6, 16, 1)
        self.layer0 = _RepVGGDeployBlock(3, 64, 2, 1)
        cin, idx = 64, 1
        for li, (w, nb) in enumerate(zip(widths, blocks), start=1):
            stage = []
            for b in range(nb):
                groups = 2 if (idx % 2 == 0 and idx <= 26) else 1
                stage.append(_RepVGGDeployBlock(cin, w, 2 if b == 0 else 1, groups))
                cin = w
                idx += 1
            setattr(self, f"layer{li}", nn.Sequential(*stage))
        self.gap = nn.AdaptiveAvgPool2d(1)
        self.linear_reg = nn.Linear(2048, 6)

    def forward(self, x):
        x = self.layer4(self.layer3(self.layer2(self.layer1(self.layer0(x)))))
        x = torch.flatten(self.gap(x), 1)
        return self.linear_reg(x)


def ortho6d_to_euler_deg(x6: torch.Tensor) -> torch.Tensor:
    """6-D -> rotation matrix -> Euler (x,y,z) in degrees (mysixdrepnet.py:272-315, 826-828)."""
    def _norm(v):
        return v / torch.sqrt(v.pow(2).sum(1)).clamp_min(1e-8)[:, None]
    x = _norm(x6[:, 0:3])
    z = _norm(torch.cross(x, x6[:, 3:6], dim=1))
    y = torch.cross(z, x, dim=1)
    R = torch.stack((x, y, z), dim=2)
    sy = torch.sqrt(R[:, 0, 0] ** 2 + R[:, 1, 0] ** 2)
    singular = (sy < 1e-6).float()
    ex = torch.atan2(R[:, 2, 1], R[:, 2, 2]) * (1 - singular) + torch.atan2(-R[:, 1, 2], R[:, 1, 1]) * singular
    ey = torch.atan2(-R[:, 2, 0], sy)
    ez = torch.atan2(R[:, 1, 0], R[:, 0, 0]) * (1 - singular)
    return torch.stack((ex, ey, ez), dim=1) * (180.0 / math.pi)


class SixDRepNet_Detector:
    """Plain-Python holder like the reference's (mysixdrepnet.py:771-833): NOT an nn.Module, so its weights stay out
    of `state_dict()` / `parameters()` exactly as in the reference.  No download: weights come from
    `load_state_dict` on `.model` (or the seeded recipe)."""

    def __init__(self, gpu_id: int = -1, dict_path: str = ""):
        self.gpu = gpu_id
        self.model = SixDRepNetBackbone().eval()
        if dict_path:
            self.model.load_state_dict(torch.load(dict_path, map_location="cpu"))

    def predict(self, img):
        x = self.model(img)
        return ortho6d_to_euler_deg(x[:, :6]), x[:, 6:]


def _fold_conv_bn(conv: nn.Conv2d, bn: nn.BatchNorm2d) -> nn.Conv2d:
    """conv -> BatchNorm(eval) as one conv (float64 fold), channels_last weights."""
    s = bn.weight.double() / torch.sqrt(bn.running_var.double() + bn.eps)
    w = conv.weight.double() * s.view(-1, 1, 1, 1)
    b0 = conv.bias.double() if conv.bias is not None else torch.zeros_like(s)
    b = (b0 - bn.running_mean.double()) * s + bn.bias.double()
    out = nn.Conv2d(conv.in_channels, conv.out_channels, conv.kernel_size, conv.stride, conv.padding,
                    conv.dilation, conv.groups, bias=True).to(conv.weight.device)
    out.weight.data = w.float().contiguous(memory_format=torch.channels_last)
    out.bias.data = b.float()
    return out.eval()


class _FoldedResNetTrunk(nn.Module):
    """Inference plan of a torchvision-style ResNet trunk: BatchNorm folded into the convolutions, channels_last,
    in-place ReLU.  Same arithmetic as the stock eval-mode modules up to fp32 rounding; built lazily from the live
    parameters and never registered (so `state_dict()` keeps the reference's keys)."""

    def __init__(self, conv1, bn1, maxpool, layers, tail=None):
        super().__init__()
        self.stem = _fold_conv_bn(conv1, bn1)
        self.maxpool = maxpool
        blocks = []
        for layer in layers:
            for blk in layer:
                kind = type(blk).__name__
                if kind == "BasicBlock":
                    convs = [_fold_conv_bn(blk.conv1, blk.bn1), _fold_conv_bn(blk.conv2, blk.bn2)]
                else:   # Bottleneck
                    convs = [_fold_conv_bn(blk.conv1, blk.bn1), _fold_conv_bn(blk.conv2, blk.bn2),
                             _fold_conv_bn(blk.conv3, blk.bn3)]
                down = None if blk.downsample is None else _fold_conv_bn(blk.downsample[0], blk.downsample[1])
                blocks.append((nn.ModuleList(convs), down))
        self.convs = nn.ModuleList([c for c, _ in blocks])
        self.downs = nn.ModuleList([d if d is not None else nn.Identity() for _, d in blocks])
        self.has_down = [d is not None for _, d in blocks]

    def forward(self, x):
        x = F.relu_(self.stem(x.contiguous(memory_format=torch.channels_last)))
        x = self.maxpool(x)
        for convs, down, hd in zip(self.convs, self.downs, self.has_down):
            idt = down(x) if hd else x
            y = x
            for i, c in enumerate(convs):
                y = c(y)
                if i + 1 < len(convs):
                    y = F.relu_(y)
            x = F.relu_(y.add_(idt))
        return x


def _versions(mod: nn.Module):
    """Per-tensor (data_ptr, version) of every parameter / buffer + device (see model._sig for the `.data` caveat)."""
    ts = list(mod.parameters()) + list(mod.buffers())
    return (tuple((t.data_ptr(), t._version) for t in ts), str(ts[0].device) if ts else "")


def _wants_grad(mod: nn.Module, x: torch.Tensor) -> bool:
    return torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in mod.parameters()))


def resnet_block_autograd(blk: nn.Module, x: torch.Tensor) -> torch.Tensor:
    """torchvision `BasicBlock` / `Bottleneck` (v1.5: stride on conv2) forward on the differentiable libmpb200 operators
    (ops.conv_train: tcgen05 forward / data gradient + tensor-core weight gradient; ops.batch_norm_train: batch statistics)."""
    from . import ops
    cb = ops.conv_bn_train                # (train mode: the batch statistics come out of the convolution's epilogue)
    idt = x
    if blk.downsample is not None:
        idt = cb(x, blk.downsample[0], blk.downsample[1])
    out = torch.relu(cb(x, blk.conv1, blk.bn1))
    out = cb(out, blk.conv2, blk.bn2)
    if hasattr(blk, "conv3"):
        out = cb(torch.relu(out), blk.conv3, blk.bn3)
    return torch.relu(out + idt)


def resnet_trunk_autograd(conv1, bn1, maxpool, layers, x: torch.Tensor) -> torch.Tensor:
    """Stem conv + BatchNorm + ReLU, the max-pool, residual stages -- the differentiable form of `ResNetTrunkPlan`."""
    from . import ops
    x = torch.relu(ops.conv_bn_train(x, conv1, bn1))
    as_int = lambda v: v if isinstance(v, int) else (v[0] if len(set(v)) == 1 else -1)
    if (as_int(maxpool.kernel_size), as_int(maxpool.stride), as_int(maxpool.padding)) == (3, 2, 1) and x.shape[1] % 4 == 0 \
            and x.shape[2] % 2 == 0 and x.shape[3] % 2 == 0 and not getattr(maxpool, "ceil_mode", False):
        x = ops.MaxPool3x3s2Function.apply(x)              # CUDA forward + backward (one byte of arg-max per output element)
    else:
        x = F.max_pool2d(x, maxpool.kernel_size, maxpool.stride, maxpool.padding)
    for layer in layers:
        for blk in layer:
            x = resnet_block_autograd(blk, x)
    return x


def _require_mpb200(mod: nn.Module, x: torch.Tensor) -> None:
    """The libmpb200 backend serves CUDA tensors in inference mode only; everything else raises (no silent fallback)."""
    if not x.is_cuda:
        raise RuntimeError(f"{type(mod).__name__}: input is on {x.device}; the B200 path has no CPU fallback "
                           "(set .backend = 'cudnn' explicitly to run the stock ATen graph)")
    if mod.training:
        raise NotImplementedError(f"{type(mod).__name__}: train-mode BatchNorm is not implemented on the B200 path "
                                  "(SURVEY.md 8f-2); call .eval(), or set .backend = 'cudnn' for the stock ATen graph")
    if torch.is_grad_enabled() and (x.requires_grad or any(p.requires_grad for p in mod.parameters())):
        raise NotImplementedError(f"{type(mod).__name__}: backward is not implemented on the B200 path (SURVEY.md "
                                  "8f-2); call under torch.no_grad(), or set .backend = 'cudnn' for the stock ATen graph")


class Emtn(nn.Module):
    """model.py:869-907.  forward(x) -> (rotations [B,3] degrees, translation [B,3], expression z [B,512])."""

    def __init__(self):
        super().__init__()
        self.head_pose_net = cifar_resnet18()
        self.head_pose_net.fc = nn.Linear(self.head_pose_net.fc.in_features, 6)
        self.rotation_net = SixDRepNet_Detector()
        model = cifar_resnet18(num_classes=512)
        self.expression_net = nn.Sequential(*list(model.children())[:-1])
        self.expression_net.adaptive_pool = nn.AdaptiveAvgPool2d(FEATURE_SIZE)
        self.fc = nn.Linear(2048, COMPRESS_DIM)
        self.backend = "mpb200"

    def _apply(self, fn, *a, **k):
        # the detector is not a registered sub-module (reference quirk); keep it on the same device/dtype anyway
        self.rotation_net.model._apply(fn, *a, **k)
        return super()._apply(fn, *a, **k)

    def _plans(self):
        sig = (_versions(self.head_pose_net), _versions(self.expression_net))
        c = self.__dict__.get("_mp_plans")
        if c is None or c[0] != sig:
            hp, ex = self.head_pose_net, self.expression_net
            with torch.no_grad():
                c = (sig, (_FoldedResNetTrunk(hp.conv1, hp.bn1, hp.maxpool, [hp.layer1, hp.layer2, hp.layer3, hp.layer4]),
                           _FoldedResNetTrunk(ex[0], ex[1], ex[3], [ex[4], ex[5], ex[6], ex[7]])))
            self.__dict__["_mp_plans"] = c
        return c[1]

    def _forward_autograd(self, x):
        """Differentiable / train-mode form (train.py:289-293 calls it on generated frames).  The two CIFAR
        ResNet-18 trunks run on the libmpb200 autograd Functions (train-mode BatchNorm = batch statistics).  6DRepNet is not a
        registered sub-module in the reference: `Gbase.train()` never reaches it and no optimizer holds its weights, so the Euler
        angles come from the inference kernels without a graph (the reference's graph through them only reaches the detector's own,
        never-updated weights and the input frame)."""
        from . import emtn_cuda
        with torch.no_grad():
            rotations = emtn_cuda.rotation_forward(self, x.detach())
        hp, ex = self.head_pose_net, self.expression_net
        f = resnet_trunk_autograd(hp.conv1, hp.bn1, hp.maxpool, (hp.layer1, hp.layer2, hp.layer3, hp.layer4), x)
        translation = hp.fc(torch.flatten(F.adaptive_avg_pool2d(f, 1), 1))[:, 3:]
        g = resnet_trunk_autograd(ex[0], ex[1], ex[3], (ex[4], ex[5], ex[6], ex[7]), x)
        g = F.adaptive_avg_pool2d(F.adaptive_avg_pool2d(g, 1), FEATURE_SIZE)      # model.py:880-881
        return rotations, translation, self.fc(torch.flatten(g, start_dim=1))

    def forward(self, x):
        if getattr(self, "backend", "mpb200") == "mpb200":
            if x.is_cuda and (self.training or _wants_grad(self, x)):
                return tuple(t.contiguous() for t in self._forward_autograd(x.float()))
            _require_mpb200(self, x)
            from . import emtn_cuda          # libmpb200 tcgen05 kernels
            with torch.no_grad():
                return emtn_cuda.emtn_forward(self, x)
        if self.training or torch.is_grad_enabled():
            # backend == "cudnn" (explicit opt-in): train-mode BatchNorm / autograd on the reference's module graph
            rotations, _ = self.rotation_net.predict(x)
            head_pose = self.head_pose_net(x)
            translation = head_pose[:, 3:]
            expression = self.fc(torch.flatten(self.expression_net(x), start_dim=1))
            return rotations, translation, expression
        # stock cuDNN plan: BatchNorm folded, channels_last (no NCHW<->NHWC transposes around the cuDNN kernels)
        hp_trunk, ex_trunk = self._plans()
        xcl = x.contiguous(memory_format=torch.channels_last)
        self.rotation_net.model.to(memory_format=torch.channels_last)
        rotations, _ = self.rotation_net.predict(xcl)
        hp = torch.flatten(F.adaptive_avg_pool2d(hp_trunk(xcl), 1), 1)
        translation = self.head_pose_net.fc(hp)[:, 3:]
        ex = F.adaptive_avg_pool2d(F.adaptive_avg_pool2d(ex_trunk(xcl), 1), FEATURE_SIZE)   # model.py:880-881
        expression = self.fc(torch.flatten(ex, start_dim=1))
        return rotations, translation, expression
