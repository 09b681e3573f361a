"""DarkNet-53 detection backbone (YOLOv3): LeakyReLU(0.1), dict in, three feature maps out.

Mirrors tlxcv/models/detection/backbones/darknet.py (``DarkNet`` :220-312).
Input is ``{"images": NCHW}`` (:300); returns the stages in ``return_idx``
(:308-309), i.e. C3/C4/C5 = 256x76x76, 512x38x38, 1024x19x19 at 608x608.
The reference keeps the stage ``Blocks`` and ``DownSample`` modules in plain
Python lists (:270-271); here, as in the oracle, they register as
``darknet_conv_block_list.{i}`` / ``downsample_list.{i}``.
"""
from __future__ import annotations

from .. import add, nn

DarkNet_cfg = {53: [1, 2, 8, 8, 4]}


class ConvBNLayer(nn.Module):
    def __init__(self, ch_in, ch_out, filter_size=3, stride=1, groups=1, padding=0, act="leaky",
                 data_format="channels_first", name="", **kwargs):
        super().__init__(name=name)
        self.conv = nn.GroupConv2d(in_channels=ch_in, out_channels=ch_out, kernel_size=filter_size, stride=stride,
                                   padding=padding, n_group=groups, b_init=False, data_format=data_format)
        self.batch_norm = nn.BatchNorm2d(num_features=ch_out, data_format=data_format)
        if act != "leaky":
            raise NotImplementedError(act)
        self.act = nn.LeakyReLU(0.1)

    def forward(self, x):
        return self.act(self.batch_norm(self.conv(x)))


class DownSample(nn.Module):
    def __init__(self, ch_in, ch_out, filter_size=3, stride=2, padding=1, data_format="channels_first", **kwargs):
        super().__init__()
        self.conv_bn_layer = ConvBNLayer(ch_in, ch_out, filter_size, stride, padding=padding,
                                         data_format=data_format)
        self.ch_out = ch_out

    def forward(self, x):
        return self.conv_bn_layer(x)


class BasicBlock(nn.Module):
    def __init__(self, ch_in, ch_out, data_format="channels_first", **kwargs):
        super().__init__()
        if ch_in != ch_out or ch_in % 2:
            raise ValueError(f"ch_in and ch_out should be the same even int, got {ch_in}, {ch_out}")
        self.conv1 = ConvBNLayer(ch_in, ch_out // 2, 1, 1, padding=0, data_format=data_format)
        self.conv2 = ConvBNLayer(ch_out // 2, ch_out, 3, 1, padding=1, data_format=data_format)

    def forward(self, x):
        return add(value=x, bias=self.conv2(self.conv1(x)))      # no activation after the add (:158)


class Blocks(nn.Module):
    def __init__(self, ch_in, ch_out, count, name=None, data_format="channels_first", **kwargs):
        super().__init__(name=name)
        self.basicblock0 = BasicBlock(ch_in, ch_out, data_format=data_format)
        self.res_blocks = nn.Sequential([BasicBlock(ch_out, ch_out, data_format=data_format)
                                         for _ in range(1, count)])
        self.ch_out = ch_out

    def forward(self, x):
        return self.res_blocks(self.basicblock0(x))


class DarkNet(nn.Module):
    def __init__(self, depth=53, freeze_at=-1, return_idx=[2, 3, 4], num_stages=5, norm_type="bn", norm_decay=0.0,
                 freeze_norm=False, data_format="channels_first"):
        super().__init__()
        self.depth, self.freeze_at, self.return_idx, self.num_stages = depth, freeze_at, list(return_idx), num_stages
        self.stages = DarkNet_cfg[depth][:num_stages]
        self.conv0 = ConvBNLayer(3, 32, 3, 1, padding=1, data_format=data_format)
        self.downsample0 = DownSample(32, 64, data_format=data_format)
        widths = [64, 128, 256, 512, 1024]
        self._out_channels = [widths[i] for i in range(len(self.stages)) if i in self.return_idx]
        self.darknet_conv_block_list = []
        self.downsample_list = []
        for i, count in enumerate(self.stages):
            self.darknet_conv_block_list.append(
                Blocks(widths[i], widths[i], count, name=f"stage.{i}", data_format=data_format))
        for i in range(num_stages - 1):
            self.downsample_list.append(DownSample(widths[i], widths[i + 1], data_format=data_format))

    def forward(self, inputs):
        out = self.downsample0(self.conv0(inputs["images"]))
        feats = []
        for i, stage in enumerate(self.darknet_conv_block_list):
            out = stage(out)
            if i in self.return_idx:
                feats.append(out)
            if i < self.num_stages - 1:
                out = self.downsample_list[i](out)
        return feats
