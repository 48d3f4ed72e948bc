"""Float model zoo used by the bench / parity harness.

Architectures follow the reference's benchmark models so that the quantized hot
path sees the layer shapes BASELINE.json names:

* NIN       — reference ``micronet/models/nin.py:42-65``
* NIN-GC    — reference ``micronet/models/nin_gc.py:62-147`` (grouped convs +
              channel shuffle, cfg [256,256,256,512,512,512,1024,1024])
* ResNet-18 — reference ``micronet/models/resnet.py:122-185`` (CIFAR stem)

They are plain ``nn.Module`` graphs (BN / ReLU / pooling / shuffle stay stock
PyTorch: out of scope per SURVEY.md §2 C5); parameter names match the
reference's ``state_dict`` so checkpoints and golden fixtures load unchanged.
"""
from __future__ import annotations

import torch
import torch.nn as nn


class Add(nn.Module):
    """residual add as a module so ``prepare`` can find it (reference
    ``micronet/base_module/op.py:5-11``)."""

    def forward(self, res, shortcut):
        return res + shortcut


def shuffle_channels(x: torch.Tensor, groups: int) -> torch.Tensor:
    n, c, h, w = x.shape
    assert c % groups == 0
    return x.view(n, groups, c // groups, h, w).transpose(1, 2).contiguous().view(n, c, h, w)


class ConvBNReLU(nn.Module):
    def __init__(self, cin, cout, kernel_size, stride=1, padding=0, groups=1,
                 channel_shuffle=0, shuffle_groups=1):
        super().__init__()
        self.channel_shuffle_flag = channel_shuffle
        self.shuffle_groups = shuffle_groups
        self.conv = nn.Conv2d(cin, cout, kernel_size, stride=stride, padding=padding, groups=groups)
        self.bn = nn.BatchNorm2d(cout)
        self.relu = nn.ReLU(inplace=True)

    def forward(self, x):
        if self.channel_shuffle_flag:
            x = shuffle_channels(x, self.shuffle_groups)
        return self.relu(self.bn(self.conv(x)))


class _SeqNet(nn.Module):
    def forward(self, x):
        x = self.model(x)
        return x.view(x.size(0), -1)


class NIN(_SeqNet):
    def __init__(self, cfg=None):
        super().__init__()
        c = cfg or [192, 160, 96, 192, 192, 192, 192, 192]
        self.model = nn.Sequential(
            ConvBNReLU(3, c[0], 5, 1, 2),
            ConvBNReLU(c[0], c[1], 1),
            ConvBNReLU(c[1], c[2], 1),
            nn.MaxPool2d(kernel_size=3, stride=2, padding=1),
            ConvBNReLU(c[2], c[3], 5, 1, 2),
            ConvBNReLU(c[3], c[4], 1),
            ConvBNReLU(c[4], c[5], 1),
            nn.MaxPool2d(kernel_size=3, stride=2, padding=1),
            ConvBNReLU(c[5], c[6], 3, 1, 1),
            ConvBNReLU(c[6], c[7], 1),
            ConvBNReLU(c[7], 10, 1),
            nn.AvgPool2d(kernel_size=8, stride=1, padding=0),
        )


class NINGC(_SeqNet):
    def __init__(self, cfg=None):
        super().__init__()
        c = cfg or [256, 256, 256, 512, 512, 512, 1024, 1024]
        self.model = nn.Sequential(
            ConvBNReLU(3, c[0], 5, 1, 2),
            ConvBNReLU(c[0], c[1], 1, groups=2, channel_shuffle=0),
            ConvBNReLU(c[1], c[2], 1, groups=2, channel_shuffle=1, shuffle_groups=2),
            nn.MaxPool2d(kernel_size=2, stride=2, padding=0),
            ConvBNReLU(c[2], c[3], 3, 1, 1, groups=16, channel_shuffle=1, shuffle_groups=2),
            ConvBNReLU(c[3], c[4], 1, groups=4, channel_shuffle=1, shuffle_groups=16),
            ConvBNReLU(c[4], c[5], 1, groups=4, channel_shuffle=1, shuffle_groups=4),
            nn.MaxPool2d(kernel_size=2, stride=2, padding=0),
            ConvBNReLU(c[5], c[6], 3, 1, 1, groups=32, channel_shuffle=1, shuffle_groups=4),
            ConvBNReLU(c[6], c[7], 1, groups=8, channel_shuffle=1, shuffle_groups=32),
            ConvBNReLU(c[7], 10, 1),
            nn.AvgPool2d(kernel_size=8, stride=1, padding=0),
        )


class BasicBlock(nn.Module):
    expansion = 1

    def __init__(self, cin, cout, stride=1):
        super().__init__()
        self.residual_function = nn.Sequential(
            nn.Conv2d(cin, cout, 3, stride=stride, padding=1, bias=False),
            nn.BatchNorm2d(cout),
            nn.ReLU(inplace=True),
            nn.Conv2d(cout, cout, 3, padding=1, bias=False),
            nn.BatchNorm2d(cout),
        )
        self.shortcut = nn.Sequential()
        if stride != 1 or cin != cout:
            self.shortcut = nn.Sequential(
                nn.Conv2d(cin, cout, 1, stride=stride, bias=False), nn.BatchNorm2d(cout))
        self.add = Add()
        self.act = nn.ReLU(inplace=True)

    def forward(self, x):
        return self.act(self.add(self.residual_function(x), self.shortcut(x)))


class ResNet(nn.Module):
    def __init__(self, blocks=(2, 2, 2, 2), num_classes=10, widths=(64, 128, 256, 512)):
        super().__init__()
        self.in_channels = widths[0]
        self.conv1 = nn.Sequential(
            nn.Conv2d(3, widths[0], 3, padding=1, bias=False), nn.BatchNorm2d(widths[0]),
            nn.ReLU(inplace=True))
        self.conv2_x = self._stage(widths[0], blocks[0], 1)
        self.conv3_x = self._stage(widths[1], blocks[1], 2)
        self.conv4_x = self._stage(widths[2], blocks[2], 2)
        self.conv5_x = self._stage(widths[3], blocks[3], 2)
        self.avg_pool = nn.AdaptiveAvgPool2d((1, 1))
        self.fc = nn.Linear(widths[3], num_classes)

    def _stage(self, cout, n, stride):
        layers = []
        for s in [stride] + [1] * (n - 1):
            layers.append(BasicBlock(self.in_channels, cout, s))
            self.in_channels = cout
        return nn.Sequential(*layers)

    def forward(self, x):
        x = self.conv5_x(self.conv4_x(self.conv3_x(self.conv2_x(self.conv1(x)))))
        x = self.avg_pool(x)
        return self.fc(x.view(x.size(0), -1))


def resnet18(**kw):
    return ResNet((2, 2, 2, 2), **kw)


def init_like_reference(model: nn.Module) -> nn.Module:
    """xavier-uniform convs, N(0, 0.01) linears, zero biases — the init the
    reference's training scripts apply (wbwtab/main.py:309-317)."""
    for m in model.modules():
        if isinstance(m, nn.Conv2d):
            nn.init.xavier_uniform_(m.weight)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
        elif isinstance(m, nn.Linear):
            nn.init.normal_(m.weight, 0, 0.01)
            if m.bias is not None:
                nn.init.zeros_(m.bias)
    return model
