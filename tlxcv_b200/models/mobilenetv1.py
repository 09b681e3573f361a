"""MobileNetV1 — the depthwise model TLXCV actually exports.

Mirrors tlxcv/models/classification/mobilenetv1.py (``MobileNetV1`` :105-262):
paths ``conv1.{0,1}``, ``dwsl.{0..12}._depthwise_conv/_pointwise_conv.{0,1}``,
``fc``; kwargs ``scale, num_classes, with_pool, data_format``.
"""
from __future__ import annotations

from .. import get_tensor_shape, nn, reshape

# (depthwise channels, pointwise out, stride) before scaling — mobilenetv1.py:133-244
_STAGES = [(32, 64, 1), (64, 128, 2), (128, 128, 1), (128, 256, 2), (256, 256, 1), (256, 512, 2),
           (512, 512, 1), (512, 512, 1), (512, 512, 1), (512, 512, 1), (512, 512, 1), (512, 1024, 2),
           (1024, 1024, 1)]


class ConvNormActivation(nn.Sequential):
    """``Sequential(conv(pad=(k-1)//2*dil), bn, act)`` (mobilenetv1.py:7-65)."""

    def __init__(self, in_channels, out_channels, kernel_size=3, stride=1, padding=None, groups=1,
                 batch_norm=nn.BatchNorm2d, activation_layer=nn.ReLU, dilation=1, bias=None,
                 data_format="channels_first"):
        if padding is None:
            padding = (kernel_size - 1) // 2 * dilation
        if bias is None:
            bias = batch_norm is None
        parts = [nn.GroupConv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size,
                                stride=stride, padding=padding, dilation=dilation, n_group=groups, b_init=bias,
                                data_format=data_format)]
        if batch_norm is not None:
            parts.append(batch_norm(num_features=out_channels, data_format=data_format))
        if activation_layer is not None:
            parts.append(activation_layer())
        super().__init__(*parts)


class DepthwiseSeparable(nn.Module):
    def __init__(self, in_channels, out_channels1, out_channels2, num_groups, stride, scale,
                 data_format="channels_first", name=None):
        super().__init__(name=name)
        mid = int(out_channels1 * scale)
        self._depthwise_conv = ConvNormActivation(in_channels, mid, 3, stride, 1, groups=int(num_groups * scale),
                                                  data_format=data_format)
        self._pointwise_conv = ConvNormActivation(mid, int(out_channels2 * scale), 1, 1, 0, data_format=data_format)

    def forward(self, x):
        return self._pointwise_conv(self._depthwise_conv(x))


class MobileNetV1(nn.Module):
    def __init__(self, scale=1.0, num_classes=1000, with_pool=True, data_format="channels_first"):
        super().__init__()
        self.scale, self.num_classes, self.with_pool = scale, num_classes, with_pool
        self.conv1 = ConvNormActivation(3, int(32 * scale), 3, 2, 1, data_format=data_format)
        names = ["conv2_1", "conv2_2", "conv3_1", "conv3_2", "conv4_1", "conv4_2"] + \
                [f"conv5_{i}" for i in range(1, 7)] + ["conv6"]
        self.dwsl = nn.Sequential(*[
            DepthwiseSeparable(int(c1 * scale), c1, c2, c1, s, scale, data_format=data_format, name=nm)
            for (c1, c2, s), nm in zip(_STAGES, names)])
        if with_pool:
            self.pool2d_avg = nn.AdaptiveAvgPool2d(1, data_format=data_format)
        if num_classes > 0:
            self.fc = nn.Linear(in_features=int(1024 * scale), out_features=num_classes)

    def forward(self, x):
        x = self.dwsl(self.conv1(x))
        if self.with_pool:
            x = self.pool2d_avg(x)
        if self.num_classes > 0:
            x = self.fc(reshape(x, (get_tensor_shape(x)[0], -1)))
        return x
