"""MobileNetV2.

Mirrors tlxcv/models/classification/mobilenetv2.py (``MobileNetV2`` :43-109,
``mobilenet_v2`` :119-148) with its helpers ops/ops_fusion.py:11-48 and
utils/common_func.py:1-16: paths ``features.{0..18}[.conv].{j}[.{0,1}]``,
``classifier.1``; kwargs ``pretrained, scale, num_classes, with_pool``.  The
reference file itself needs ``paddle``/``paddle2tlx`` to import and is not
exported from ``tlxcv.models``; this one has no such dependency.
"""
from __future__ import annotations

from .. import flatten, nn
from .mobilenetv1 import ConvNormActivation

# expansion t, channels c, repeats n, first stride s — mobilenetv2.py:76-78
_SETTING = [(1, 16, 1, 1), (6, 24, 2, 2), (6, 32, 3, 2), (6, 64, 4, 2), (6, 96, 3, 1), (6, 160, 3, 2), (6, 320, 1, 1)]


def _make_divisible(v, divisor=8, min_value=None):
    floor = divisor if min_value is None else min_value
    rounded = max(floor, int(v + divisor / 2) // divisor * divisor)
    return rounded + divisor if rounded < 0.9 * v else rounded


class InvertedResidual(nn.Module):
    def __init__(self, inp, oup, stride, expand_ratio, batch_norm=nn.BatchNorm2d):
        super().__init__()
        if stride not in (1, 2):
            raise ValueError("stride must be 1 or 2")
        self.stride = stride
        hidden = int(round(inp * expand_ratio))
        self.use_res_connect = stride == 1 and inp == oup
        chain = []
        if expand_ratio != 1:
            chain.append(ConvNormActivation(inp, hidden, kernel_size=1, batch_norm=batch_norm,
                                            activation_layer=nn.ReLU6))
        chain.append(ConvNormActivation(hidden, hidden, stride=stride, groups=hidden, batch_norm=batch_norm,
                                        activation_layer=nn.ReLU6))
        chain.append(nn.GroupConv2d(in_channels=hidden, out_channels=oup, kernel_size=1, stride=1, padding=0,
                                    b_init=(), data_format="channels_first"))
        chain.append(batch_norm(num_features=oup, data_format="channels_first"))
        self.conv = nn.Sequential(chain)

    def forward(self, x):
        y = self.conv(x)
        return x + y if self.use_res_connect else y


class MobileNetV2(nn.Module):
    def __init__(self, scale=1.0, num_classes=1000, with_pool=True):
        super().__init__()
        self.num_classes, self.with_pool = num_classes, with_pool
        cin = _make_divisible(32 * scale, 8)
        self.last_channel = _make_divisible(1280 * max(1.0, scale), 8)
        feats = [ConvNormActivation(3, cin, stride=2, activation_layer=nn.ReLU6)]
        for t, c, n, s in _SETTING:
            cout = _make_divisible(c * scale, 8)
            for i in range(n):
                feats.append(InvertedResidual(cin, cout, s if i == 0 else 1, expand_ratio=t))
                cin = cout
        feats.append(ConvNormActivation(cin, self.last_channel, kernel_size=1, activation_layer=nn.ReLU6))
        self.features = nn.Sequential(feats)
        if with_pool:
            self.pool2d_avg = nn.AdaptiveAvgPool2d(1, data_format="channels_first")
        if num_classes > 0:
            self.classifier = nn.Sequential([nn.Dropout(0.2),
                                             nn.Linear(in_features=self.last_channel, out_features=num_classes)])

    def forward(self, x):
        x = self.features(x)
        if self.with_pool:
            x = self.pool2d_avg(x)
        if self.num_classes > 0:
            x = self.classifier(flatten(x, 1))
        return x


def mobilenet_v2(pretrained=False, scale=1.0, **kwargs):
    if pretrained:
        raise NotImplementedError("pretrained weights need paddle2tlx + network (mobilenetv2.py:112-116)")
    return MobileNetV2(scale=scale, **kwargs)
