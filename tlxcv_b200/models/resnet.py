"""ResNet-18/34/50/101/152 and wide-50/101 (v1.5: the stride sits on the 3x3).

Mirrors tlxcv/models/classification/resnet.py: constructors :308-382, ``ResNet``
:159-300.  Module paths ``conv1, bn1, layer{1-4}.{i}.conv{1-3}/bn{1-3}/
downsample.{0,1}, fc`` and kwargs ``width, num_classes, with_pool, groups,
data_format, name`` are the reference's.
"""
from __future__ import annotations

from .. import FlattenReshape, nn

_DEPTHS = {18: (2, 2, 2, 2), 34: (3, 4, 6, 3), 50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}


def _conv(cin, cout, k, stride=1, pad=0, groups=1, dilation=1, fmt="channels_first"):
    return nn.GroupConv2d(in_channels=cin, out_channels=cout, kernel_size=k, stride=stride, padding=pad,
                          dilation=dilation, n_group=groups, b_init=(), data_format=fmt)


def _bn(c, fmt):
    return nn.BatchNorm2d(num_features=c, data_format=fmt)


class _Residual(nn.Module):
    """Shared tail of both block types: ``relu(branch(x) + shortcut(x))`` (resnet.py:73-77,152-156)."""

    expansion = 1

    def _shortcut(self, x):
        return x if self.downsample is None else self.downsample(x)


class BasicBlock(_Residual):
    expansion = 1

    def __init__(self, cin, planes, stride=1, downsample=None, groups=1, base_width=64, dilation=1,
                 data_format="channels_first"):
        super().__init__()
        if dilation > 1:
            raise NotImplementedError("BasicBlock does not dilate (resnet.py:34-36)")
        self.conv1, self.bn1 = _conv(cin, planes, 3, stride, 1, fmt=data_format), _bn(planes, data_format)
        self.relu = nn.ReLU()
        self.conv2, self.bn2 = _conv(planes, planes, 3, 1, 1, fmt=data_format), _bn(planes, data_format)
        self.downsample, self.stride = downsample, stride

    def forward(self, x):
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.bn2(self.conv2(y))
        return self.relu(y + self._shortcut(x))


class BottleneckBlock(_Residual):
    expansion = 4

    def __init__(self, cin, planes, stride=1, downsample=None, groups=1, base_width=64, dilation=1,
                 data_format="channels_first"):
        super().__init__()
        mid = int(planes * (base_width / 64.0)) * groups           # resnet.py:98
        self.conv1, self.bn1 = _conv(cin, mid, 1, fmt=data_format), _bn(mid, data_format)
        self.conv2 = _conv(mid, mid, 3, stride, dilation, groups, dilation, fmt=data_format)
        self.bn2 = _bn(mid, data_format)
        self.conv3, self.bn3 = _conv(mid, planes * 4, 1, fmt=data_format), _bn(planes * 4, data_format)
        self.relu = nn.ReLU()
        self.downsample, self.stride = downsample, stride

    def forward(self, x):
        y = self.relu(self.bn1(self.conv1(x)))
        y = self.relu(self.bn2(self.conv2(y)))
        y = self.bn3(self.conv3(y))
        return self.relu(y + self._shortcut(x))


class ResNet(nn.Module):
    def __init__(self, block, depth=50, width=64, num_classes=1000, with_pool=True, groups=1,
                 data_format="channels_first", name=None):
        super().__init__(name=name)
        self.groups, self.base_width = groups, width
        self.num_classes, self.with_pool = num_classes, with_pool
        self.in_channels, self.dilation = 64, 1
        self.conv1, self.bn1 = _conv(3, 64, 7, 2, 3, fmt=data_format), _bn(64, data_format)
        self.relu = nn.ReLU()
        self.maxpool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1, data_format=data_format)
        for i, (planes, count) in enumerate(zip((64, 128, 256, 512), _DEPTHS[depth])):
            stage = self._make_layer(block, planes, count, stride=1 if i == 0 else 2, data_format=data_format)
            setattr(self, f"layer{i + 1}", stage)
        if with_pool:
            self.avgpool = nn.AdaptiveAvgPool2d((1, 1), data_format=data_format)
        self.flatten = FlattenReshape()
        if num_classes > 0:
            self.fc = nn.Linear(in_features=512 * block.expansion, out_features=num_classes)

    def _make_layer(self, block, planes, count, stride=1, dilate=False, data_format="channels_first"):
        prev_dilation = self.dilation
        if dilate:
            self.dilation, stride = self.dilation * stride, 1
        cout = planes * block.expansion
        shortcut = None
        if stride != 1 or self.in_channels != cout:                  # resnet.py:246
            shortcut = nn.Sequential([_conv(self.in_channels, cout, 1, stride, fmt=data_format),
                                      _bn(cout, data_format)])
        blocks = [block(self.in_channels, planes, stride, shortcut, self.groups, self.base_width, prev_dilation,
                        data_format=data_format)]
        self.in_channels = cout
        blocks += [block(cout, planes, groups=self.groups, base_width=self.base_width, data_format=data_format)
                   for _ in range(count - 1)]
        return nn.Sequential(blocks)

    def forward(self, x):
        x = self.maxpool(self.relu(self.bn1(self.conv1(x))))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        if self.with_pool:
            x = self.avgpool(x)
        if self.num_classes > 0:
            x = self.fc(self.flatten(x))
        return x


def _build(arch, block, depth, pretrained, **kwargs):
    if pretrained:
        raise NotImplementedError("pretrained weights are not available offline (the reference ignores the flag, "
                                  "resnet.py:303-305)")
    return ResNet(block, depth, name=arch, **kwargs)


def resnet18(pretrained=False, **kw):
    return _build("resnet18", BasicBlock, 18, pretrained, **kw)


def resnet34(pretrained=False, **kw):
    return _build("resnet34", BasicBlock, 34, pretrained, **kw)


def resnet50(pretrained=False, **kw):
    return _build("resnet50", BottleneckBlock, 50, pretrained, **kw)


def resnet101(pretrained=False, **kw):
    return _build("resnet101", BottleneckBlock, 101, pretrained, **kw)


def resnet152(pretrained=False, **kw):
    return _build("resnet152", BottleneckBlock, 152, pretrained, **kw)


def wide_resnet50_2(pretrained=False, **kw):
    return _build("wide_resnet50_2", BottleneckBlock, 50, pretrained, width=128, **kw)


def wide_resnet101_2(pretrained=False, **kw):
    return _build("wide_resnet101_2", BottleneckBlock, 101, pretrained, width=128, **kw)
