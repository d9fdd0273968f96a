"""Motion encoder `Emtn` and the `CustomResNet50` descriptor branch as stock PyTorch modules.

SURVEY.md section 8: `Emtn` (model.py:869-907) is on Gbase's call path but is NOT one of the hot-path rows this
round (it is row f-1, "next"); `CustomResNet50` (model.py:136-173) is a secondary row that "may stay on cuDNN at
first".  Both therefore run as ordinary torch/cuDNN modules here, with the reference's attribute names and
state_dict keys, so that `Gbase.forward` is complete and checkpoints load with strict=True.
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

    def forward(self, x):
        x = F.relu(self.bn1(self.conv1(x)))
        x = self.maxpool(x)
        x = self.layer3(self.layer2(self.layer1(x)))
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

    def _apply(self, fn, *a, **k):
        # the detector is not a registered sub-module (reference quirk); keep it on the same device/dtype anyway
        self.rotation_net.model._apply(fn, *a, **k)
        return super()._apply(fn, *a, **k)

    def forward(self, x):
        rotations, _ = self.rotation_net.predict(x)
        head_pose = self.head_pose_net(x)
        translation = head_pose[:, 3:]
        expression = self.fc(torch.flatten(self.expression_net(x), start_dim=1))
        return rotations, translation, expression
