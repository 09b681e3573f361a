"""DarkNet-53 classifier (ReLU, global pool + FC).

Mirrors tlxcv/models/classification/darknet53.py (``DarkNet`` :56-133,
``darknet53`` :143): paths ``_conv1, _conv2, _basic_block_XY._conv{1,2}._conv/_bn,
_downsample_N, _out``; kwarg ``class_num``.  BN carries the ReLU (``act='relu'``).
"""
from __future__ import annotations

from .. import add, nn, ops
from ..nn.initializers import xavier_uniform

_STAGES = (1, 2, 8, 8, 4)


class ConvBNLayer(nn.Module):
    def __init__(self, input_channels, output_channels, filter_size, stride, padding, name=None):
        super().__init__()
        self._conv = nn.GroupConv2d(in_channels=input_channels, out_channels=output_channels,
                                    kernel_size=filter_size, stride=stride, padding=padding,
                                    W_init=xavier_uniform(), b_init=(), data_format="channels_first")
        self._bn = nn.BatchNorm(act="relu", num_features=output_channels, data_format="channels_first")

    def forward(self, x):
        return self._bn(self._conv(x))


class BasicBlock(nn.Module):
    def __init__(self, input_channels, output_channels, name=None):
        super().__init__()
        self._conv1 = ConvBNLayer(input_channels, output_channels, 1, 1, 0, name=f"{name}.0")
        self._conv2 = ConvBNLayer(output_channels, output_channels * 2, 3, 1, 1, name=f"{name}.1")

    def forward(self, x):
        return add(value=x, bias=self._conv2(self._conv1(x)))


class DarkNet53(nn.Module):
    def __init__(self, class_num=1000):
        super().__init__()
        self.stages = list(_STAGES)
        self._conv1 = ConvBNLayer(3, 32, 3, 1, 1, name="yolo_input")
        self._conv2 = ConvBNLayer(32, 64, 3, 2, 1, name="yolo_input.downsample")
        self._order = []
        width = 64
        for si, count in enumerate(_STAGES):
            for bi in range(1, count + 1):
                key = f"_basic_block_{si}{bi}"
                setattr(self, key, BasicBlock(width, width // 2, name=f"stage.{si}.{bi - 1}"))
                self._order.append(key)
            if si < len(_STAGES) - 1:
                key = f"_downsample_{si}"
                setattr(self, key, ConvBNLayer(width, width * 2, 3, 2, 1, name=f"stage.{si}.downsample"))
                self._order.append(key)
                width *= 2
        self._pool = nn.AdaptiveAvgPool2d(1, data_format="channels_first")
        self._out = nn.Linear(in_features=1024, out_features=class_num, b_init=xavier_uniform())

    def forward(self, x):
        x = self._conv2(self._conv1(x))
        for key in self._order:
            x = getattr(self, key)(x)
        x = ops.squeeze(self._pool(x), axis=[2, 3])
        return self._out(x)


def darknet53(pretrained=False, **kwargs):
    if pretrained:
        raise NotImplementedError("pretrained weights need paddle2tlx + network (darknet53.py:136-140)")
    return DarkNet53(**kwargs)
