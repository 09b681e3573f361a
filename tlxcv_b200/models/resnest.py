"""ResNeSt (split-attention ResNet) on the B200 path.

Mirrors tlxcv/models/classification/resnest.py: ``ConvBNLayer`` (:12-50, ``_conv`` + ``batch_norm`` with the activation inside
the BatchNorm), ``SplatConv`` (:84-167: a radix-major grouped 3x3 conv, global average pool of the summed radix groups, two
1x1 convs, ``rSoftmax`` (:53-82) over the radix axis, attention-weighted sum of the groups), ``BottleneckBlock`` (:170-328:
1x1 -> SplatConv -> [3x3 / stride-2 / pad-1 average pool, "avd"] -> 1x1, shortcut = [AvgPool2d(stride)] + 1x1 conv + BN,
"avg_down"), ``ResNeStLayer`` (:330-438, blocks registered as ``<layer>_bottleneck_<i>``) and ``ResNeSt`` (:441-683: three
3x3 stem convs, max-pool, four stages, global pool, ``out`` Linear).  Parameter names and order equal the reference's, so
its state dicts load unchanged.

On the kernel side the split attention is three launches besides the convs: the global average pool runs over ALL radix
groups (the reference's ``add_n`` of the groups moves into the first 1x1 conv, whose filters are repeated along C_in:
``planner.DerivedConv``), and ``TLXCV_OP_SPLAT_APPLY`` does the radix softmax, the broadcast multiply and the sum in one pass.
The radix-major 3x3 (groups = cardinality * radix, C -> radix * C) runs as a dense conv with block-diagonal filters.
Only ``radix > 1`` is on the hot path (``resnest50``, ``resnest101``; the radix-1 "fast" variant gates with a sigmoid).
"""
from __future__ import annotations

from collections import OrderedDict

from .. import add, nn, relu, reshape

__all__ = ["ResNeSt", "resnest50", "resnest101"]


class ConvBNLayer(nn.Module):
    def __init__(self, num_channels, num_filters, filter_size, stride=1, dilation=1, groups=1, act=None,
                 data_format="channels_first", name=None):
        super().__init__(name)
        self._conv = nn.GroupConv2d(in_channels=num_channels, out_channels=num_filters, kernel_size=filter_size, stride=stride,
                                    padding=(filter_size - 1) // 2, dilation=dilation, b_init=None, n_group=groups,
                                    data_format=data_format)
        self.batch_norm = nn.BatchNorm(act=act, num_features=num_filters, data_format=data_format)

    def forward(self, x):
        return self.batch_norm(self._conv(x))


class rSoftmax(nn.Module):
    """Holds radix / cardinality; the softmax itself is part of ``TLXCV_OP_SPLAT_APPLY``."""

    def __init__(self, radix, cardinality, data_format="channels_first"):
        super().__init__()
        self.radix, self.cardinality, self.data_format = radix, cardinality, data_format


class SplatConv(nn.Module):
    def __init__(self, in_channels, channels, kernel_size, stride=1, padding=0, dilation=1, groups=1, bias=True, radix=2,
                 reduction_factor=4, rectify_avg=False, data_format="channels_first", name=None):
        super().__init__(name)
        self.radix = radix
        self.conv1 = ConvBNLayer(in_channels, channels * radix, kernel_size, stride=stride, groups=groups * radix, act="relu",
                                 data_format=data_format)
        self.avg_pool2d = nn.AdaptiveAvgPool2d(1, data_format=data_format)
        inter_channels = int(max(in_channels * radix // reduction_factor, 32))
        self.conv2 = ConvBNLayer(channels, inter_channels, 1, groups=groups, act="relu", data_format=data_format)
        self.conv3 = nn.GroupConv2d(in_channels=inter_channels, out_channels=channels * radix, kernel_size=1, stride=1, padding=0,
                                    b_init=None, n_group=groups, data_format=data_format)
        self.rsoftmax = rSoftmax(radix, groups, data_format)

    def forward(self, x):
        return nn.split_attention(self, x)


class BottleneckBlock(nn.Module):
    def __init__(self, inplanes, planes, stride=1, radix=1, cardinality=1, bottleneck_width=64, avd=False, avd_first=False,
                 dilation=1, is_first=False, avg_down=False, data_format="channels_first", name=None):
        super().__init__(name)
        self.stride, self.avd, self.avd_first, self.is_first, self.avg_down = stride, avd, avd_first, is_first, avg_down
        width = int(planes * (bottleneck_width / 64.0)) * cardinality
        pooled = avd and (stride > 1 or is_first)
        self.conv1 = ConvBNLayer(inplanes, width, 1, act="relu", data_format=data_format)
        if pooled and avd_first:
            self.avg_pool2d_1 = nn.AvgPool2d(kernel_size=3, stride=stride, padding=1, data_format=data_format)
        if radix < 1:
            raise NotImplementedError("ResNeSt without a SplatConv (radix 0) is the plain ResNeXt block: use models.resnext")
        self.conv2 = SplatConv(width, width, 3, stride=1, padding=dilation, dilation=dilation, groups=cardinality, bias=False,
                               radix=radix, data_format=data_format)
        if pooled and not avd_first:
            self.avg_pool2d_2 = nn.AvgPool2d(kernel_size=3, stride=stride, padding=1, data_format=data_format)
        self.conv3 = ConvBNLayer(width, planes * 4, 1, data_format=data_format)
        self.projects = stride != 1 or inplanes != planes * 4
        if self.projects:
            if avg_down:
                self.avg_pool2d_3 = nn.AvgPool2d(kernel_size=stride if dilation == 1 else 1, stride=stride if dilation == 1 else 1,
                                                 padding=0, data_format=data_format)
            self.conv4 = nn.GroupConv2d(in_channels=inplanes, out_channels=planes * 4, kernel_size=1,
                                        stride=1 if avg_down else stride, padding=0, b_init=None, n_group=1, data_format=data_format)
            self.batch_norm = nn.BatchNorm(num_features=planes * 4, data_format=data_format)

    def forward(self, x):
        short = x
        pooled = self.avd and (self.stride > 1 or self.is_first)
        x = self.conv1(x)
        if pooled and self.avd_first:
            x = self.avg_pool2d_1(x)
        x = self.conv2(x)
        if pooled and not self.avd_first:
            x = self.avg_pool2d_2(x)
        x = self.conv3(x)
        if self.projects:
            if self.avg_down:
                short = self.avg_pool2d_3(short)
            short = self.batch_norm(self.conv4(short))
        return relu(add(short, x))


class ResNeStLayer(nn.Module):
    def __init__(self, inplanes, planes, blocks, radix, cardinality, bottleneck_width, avg_down, avd, avd_first, stride=1,
                 dilation=1, is_first=True, data_format="channels_first", name=None):
        super().__init__(name)
        if dilation not in (1, 2, 4):
            raise RuntimeError("=>unknown dilation size")
        common = dict(radix=radix, cardinality=cardinality, bottleneck_width=bottleneck_width, avg_down=avg_down, avd=avd,
                      avd_first=avd_first, data_format=data_format)
        self.bottleneck_block_list = []
        for i in range(blocks):
            if i == 0:
                block = BottleneckBlock(inplanes, planes, stride=stride, dilation=2 if dilation == 4 else 1, is_first=is_first,
                                        **common)
            else:
                block = BottleneckBlock(planes * 4, planes, dilation=dilation, **common)
            setattr(self, f"{name}_bottleneck_{i}", block)
            self.bottleneck_block_list.append(block)

    def forward(self, x):
        for block in self.bottleneck_block_list:
            x = block(x)
        return x


class ResNeSt(nn.Module):
    def __init__(self, layers=(3, 4, 6, 3), radix=2, groups=1, bottleneck_width=64, dilated=False, dilation=1, deep_stem=True,
                 stem_width=32, avg_down=True, avd=True, avd_first=False, final_drop=0.0, num_classes=1000,
                 data_format="channels_first", name=None):
        super().__init__(name)
        if final_drop:
            raise NotImplementedError("dropout in front of the classifier is a training-time layer")
        self.out_channels = 2048
        if deep_stem:
            self.stem = nn.Sequential(OrderedDict([
                ("conv1", ConvBNLayer(3, stem_width, 3, stride=2, act="relu", data_format=data_format)),
                ("conv2", ConvBNLayer(stem_width, stem_width, 3, act="relu", data_format=data_format)),
                ("conv3", ConvBNLayer(stem_width, stem_width * 2, 3, act="relu", data_format=data_format))]))
        else:
            self.stem = ConvBNLayer(3, stem_width, 7, stride=2, act="relu", data_format=data_format)
        self.max_pool2d = nn.MaxPool2d(kernel_size=3, stride=2, padding=1, data_format=data_format)
        common = dict(radix=radix, cardinality=groups, bottleneck_width=bottleneck_width, avg_down=avg_down, avd=avd,
                      avd_first=avd_first, data_format=data_format)
        if dilated or dilation == 4:
            strides, dilations = (1, 2, 1, 1), (1, 1, 2, 4)
        elif dilation == 2:
            strides, dilations = (1, 2, 2, 1), (1, 1, 1, 2)
        else:
            strides, dilations = (1, 2, 2, 2), (1, 1, 1, 1)
        inplanes = stem_width * 2 if deep_stem else stem_width
        for i, (planes, n) in enumerate(zip((64, 128, 256, 512), layers)):
            setattr(self, f"layer{i + 1}", ResNeStLayer(inplanes, planes, n, stride=strides[i], dilation=dilations[i],
                                                        is_first=(i != 0), name=f"layer{i + 1}", **common))
            inplanes = planes * 4
        self.pool2d_avg = nn.AdaptiveAvgPool2d(1, data_format=data_format)
        self.out = nn.Linear(in_features=self.out_channels, out_features=num_classes)

    def forward(self, x):
        x = self.max_pool2d(self.stem(x))
        x = self.layer4(self.layer3(self.layer2(self.layer1(x))))
        return self.out(reshape(self.pool2d_avg(x), shape=[-1, self.out_channels]))


def resnest50(pretrained=False, **kwargs):
    return ResNeSt(layers=(3, 4, 6, 3), stem_width=32, **kwargs)


def resnest101(pretrained=False, **kwargs):
    return ResNeSt(layers=(3, 4, 23, 3), stem_width=64, **kwargs)
