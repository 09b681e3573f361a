"""MobileNet-v1 detection backbone (SSD / YOLOv3-MobileNet) on the B200 path.

Mirrors tlxcv/models/detection/backbones/mobilenet_v1.py: ``ConvBNLayer`` (:8-50, ``_conv`` + ``my_batch_norm`` + ReLU /
ReLU6), ``DepthwiseSeparable`` (:53-103), ``ExtraBlock`` (:106-154, 1x1 then 3x3 stride-2, ReLU6) and ``MobileNet``
(:157-245: dict input ``{"images": NCHW}``, the maps after blocks ``feature_maps`` = [4, 6, 13] come back).  ``dwsl`` and
``extra_blocks`` are plain Python lists in the reference; they register as ``dwsl.{i}`` / ``extra_blocks.{i}``.
Same kernels as the classification MobileNetV1: depthwise on the CUDA-core / slab path, pointwise on tcgen05.
"""
from __future__ import annotations

from numbers import Integral

from .. import nn

__all__ = ["MobileNet"]


class ConvBNLayer(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride, padding, num_groups=1, act="relu", conv_lr=1.0,
                 conv_decay=0.0, norm_decay=0.0, norm_type="bn", name=None, data_format="channels_first"):
        super().__init__()
        self.act = act
        self._conv = nn.GroupConv2d(kernel_size=kernel_size, stride=stride, padding=padding, in_channels=in_channels,
                                    out_channels=out_channels, W_init=nn.initializers.xavier_uniform(), b_init=False,
                                    n_group=num_groups, data_format=data_format)
        if norm_type not in ("sync_bn", "bn"):
            raise NotImplementedError(norm_type)
        self.my_batch_norm = nn.BatchNorm2d(num_features=out_channels, data_format=data_format)
        self._act = {"relu": nn.ReLU, "relu6": nn.ReLU6}[act]() if act in ("relu", "relu6") else None

    def forward(self, x):
        x = self.my_batch_norm(self._conv(x))
        return self._act(x) if self._act is not None else x


class DepthwiseSeparable(nn.Module):
    def __init__(self, in_channels, out_channels1, out_channels2, num_groups, stride, scale, name=None,
                 data_format="channels_first", **kwds):
        super().__init__()
        self._depthwise_conv = ConvBNLayer(in_channels, int(out_channels1 * scale), kernel_size=3, stride=stride, padding=1,
                                           num_groups=int(num_groups * scale), data_format=data_format)
        self._pointwise_conv = ConvBNLayer(int(out_channels1 * scale), int(out_channels2 * scale), kernel_size=1, stride=1,
                                           padding=0, data_format=data_format)

    def forward(self, x):
        return self._pointwise_conv(self._depthwise_conv(x))


class ExtraBlock(nn.Module):
    def __init__(self, in_channels, out_channels1, out_channels2, num_groups=1, stride=2, name=None,
                 data_format="channels_first", **kwds):
        super().__init__()
        self.pointwise_conv = ConvBNLayer(in_channels, int(out_channels1), kernel_size=1, stride=1, padding=0,
                                          num_groups=int(num_groups), act="relu6", data_format=data_format)
        self.normal_conv = ConvBNLayer(int(out_channels1), int(out_channels2), kernel_size=3, stride=stride, padding=1,
                                       num_groups=int(num_groups), act="relu6", data_format=data_format)

    def forward(self, x):
        return self.normal_conv(self.pointwise_conv(x))


class MobileNet(nn.Module):
    def __init__(self, norm_type="bn", norm_decay=0.0, conv_decay=0.0, scale=1, conv_learning_rate=1.0,
                 feature_maps=(4, 6, 13), with_extra_blocks=False,
                 extra_block_filters=((256, 512), (128, 256), (128, 256), (64, 128)), data_format="channels_first"):
        super().__init__()
        self.feature_maps = [feature_maps] if isinstance(feature_maps, Integral) else list(feature_maps)
        self.with_extra_blocks, self.extra_block_filters = with_extra_blocks, [list(f) for f in extra_block_filters]
        self._out_channels = []
        self.conv1 = ConvBNLayer(3, int(32 * scale), kernel_size=3, stride=2, padding=1, data_format=data_format)
        self.cfgs = [[32, 64, 1], [64, 128, 2], [128, 128, 1], [128, 256, 2], [256, 256, 1], [256, 512, 2],
                     *[[512, 512, 1] for _ in range(5)], [512, 1024, 2], [1024, 1024, 1]]           # :199-209
        self.dwsl = []
        for i, o, s in self.cfgs:
            self.dwsl.append(DepthwiseSeparable(int(i * scale), i, o, i, s, scale, data_format=data_format))
            if len(self.dwsl) in self.feature_maps:
                self._out_channels.append(int(o * scale))
        self.extra_blocks = []
        if with_extra_blocks:
            for k, (out0, out1) in enumerate(self.extra_block_filters):
                in_c = 1024 if k == 0 else self.extra_block_filters[k - 1][1]
                self.extra_blocks.append(ExtraBlock(in_c, out0, out1, data_format=data_format))
                if len(self.dwsl) + k + 1 in self.feature_maps:
                    self._out_channels.append(out1)

    def forward(self, inputs):
        outs = []
        y = self.conv1(inputs["images"])
        for idx, block in enumerate(list(self.dwsl) + list(self.extra_blocks), start=1):
            y = block(y)
            if idx in self.feature_maps:
                outs.append(y)
        return outs
