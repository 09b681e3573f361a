"""ResNeXt-50/101/152 x {32,64}x4d.

Mirrors tlxcv/models/classification/resnext.py (``ResNeXt`` :122-209, constructors
:221-242): paths ``conv._conv/.batch_norm``, ``bb_{stage}_{i}.conv{0,1,2}``,
``.short``, ``out``; kwargs ``num_classes, input_image_channel, data_format, name``.
The 3x3 is grouped (g = cardinality); BN carries the ReLU via ``act='relu'``.
"""
from __future__ import annotations

from .. import add, nn, relu, reshape
from ..nn.initializers import xavier_uniform

_DEPTHS = {50: (3, 4, 6, 3), 101: (3, 4, 23, 3), 152: (3, 8, 36, 3)}


class ConvBNLayer(nn.Module):
    def __init__(self, num_channels, num_filters, filter_size, stride=1, groups=1, act=None, name=None,
                 data_format="channels_first"):
        super().__init__(name)
        self._conv = nn.GroupConv2d(in_channels=num_channels, out_channels=num_filters, kernel_size=filter_size,
                                    stride=stride, padding=(filter_size - 1) // 2, n_group=groups, b_init=(),
                                    W_init=xavier_uniform(), data_format=data_format)
        self.batch_norm = nn.BatchNorm(act=act, num_features=num_filters, data_format=data_format)

    def forward(self, x):
        return self.batch_norm(self._conv(x))


class BottleneckBlock(nn.Module):
    def __init__(self, num_channels, num_filters, stride, cardinality, shortcut=True, name=None,
                 data_format="channels_first"):
        super().__init__(name)
        cout = num_filters * 2 if cardinality == 32 else num_filters       # resnext.py:92
        kw = dict(data_format=data_format)
        self.conv0 = ConvBNLayer(num_channels, num_filters, 1, act="relu", name=f"{name}_branch2a", **kw)
        self.conv1 = ConvBNLayer(num_filters, num_filters, 3, stride, cardinality, "relu", f"{name}_branch2b", **kw)
        self.conv2 = ConvBNLayer(num_filters, cout, 1, name=f"{name}_branch2c", **kw)
        if not shortcut:
            self.short = ConvBNLayer(num_channels, cout, 1, stride, name=f"{name}_branch1", **kw)
        self.shortcut = shortcut

    def forward(self, x):
        y = self.conv2(self.conv1(self.conv0(x)))
        s = x if self.shortcut else self.short(x)
        return relu(add(value=s, bias=y))


class ResNeXt(nn.Module):
    def __init__(self, layers=50, num_classes=1000, cardinality=32, input_image_channel=3, name=None,
                 data_format="channels_first"):
        super().__init__(name)
        if layers not in _DEPTHS:
            raise ValueError(f"supported layers are {sorted(_DEPTHS)} but input layer is {layers}")
        if cardinality not in (32, 64):
            raise ValueError(f"supported cardinality is [32, 64] but input cardinality is {cardinality}")
        self.layers, self.cardinality = layers, cardinality
        stage_in = (64, 256, 512, 1024)
        widths = (128, 256, 512, 1024) if cardinality == 32 else (256, 512, 1024, 2048)
        self.conv = ConvBNLayer(input_image_channel, 64, 7, 2, act="relu", name="res_conv1", data_format=data_format)
        self.pool2d_max = nn.MaxPool2d(kernel_size=3, stride=2, padding=1, data_format=data_format)
        self.block_list = []
        for si, count in enumerate(_DEPTHS[layers]):
            for bi in range(count):
                if layers in (101, 152) and si == 2:
                    tag = f"res{si + 2}a" if bi == 0 else f"res{si + 2}b{bi}"
                else:
                    tag = f"res{si + 2}{chr(97 + bi)}"
                cin = stage_in[si] if bi == 0 else widths[si] * (64 // cardinality)
                blk = BottleneckBlock(cin, widths[si], 2 if (bi == 0 and si != 0) else 1, cardinality,
                                      shortcut=bi != 0, name=tag, data_format=data_format)
                setattr(self, f"bb_{si}_{bi}", blk)
                self.block_list.append(blk)
        self.pool2d_avg = nn.AdaptiveAvgPool2d(1, data_format=data_format)
        self.pool2d_avg_channels = stage_in[-1] * 2
        self.out = nn.Linear(in_features=self.pool2d_avg_channels, out_features=num_classes, b_init=xavier_uniform())

    def forward(self, x):
        y = self.pool2d_max(self.conv(x))
        for blk in self.block_list:
            y = blk(y)
        y = reshape(self.pool2d_avg(y), shape=[-1, self.pool2d_avg_channels])
        return self.out(y)


def _build(layers, cardinality, pretrained, **kw):
    if pretrained:
        raise NotImplementedError("pretrained weights are not available offline")
    return ResNeXt(layers=layers, cardinality=cardinality, **kw)


def resnext50_32x4d(pretrained=False, **kw):
    return _build(50, 32, pretrained, **kw)


def resnext50_64x4d(pretrained=False, **kw):
    return _build(50, 64, pretrained, **kw)


def resnext101_32x4d(pretrained=False, **kw):
    return _build(101, 32, pretrained, **kw)


def resnext101_64x4d(pretrained=False, **kw):
    return _build(101, 64, pretrained, **kw)


def resnext152_32x4d(pretrained=False, **kw):
    return _build(152, 32, pretrained, **kw)


def resnext152_64x4d(pretrained=False, **kw):
    return _build(152, 64, pretrained, **kw)
