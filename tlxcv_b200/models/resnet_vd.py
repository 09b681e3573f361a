"""ResNet_vd segmentation backbone on the B200 path.

Mirrors tlxcv/models/segmentation/backbones/resnet_vd.py: ``ConvBNLayer`` (:7-50: optional 2x2 / stride-2 average pool in
front of the conv - the "vd" shortcut -, ``_conv`` + ``batch_norm`` + activation; dilated 3x3 convs take ``padding =
dilation``), ``BottleneckBlock`` (:53-113), ``BasicBlock`` (:116-169) and ``ResNet_vd`` (:172-326: three 3x3 stem convs,
max-pool, four stages; ``output_stride`` 8 / 16 replaces the stride of the last stages by dilation 2 / 4; returns the four
stage outputs).  ``stage_list`` is a list of lists in the reference; its blocks register as ``stage_list.{stage}.{block}``.
New on the kernel side: ``AvgPool2d`` (``TLXCV_OP_AVGPOOL``); the dilated convs run on the im2col-mode TMA path unchanged.
"""
from __future__ import annotations

from .. import add, nn

__all__ = ["ResNet_vd"]


class ConvBNLayer(nn.Module):
    def __init__(self, in_channels, out_channels, kernel_size, stride=1, dilation=1, groups=1, is_vd_mode=False, act=None,
                 data_format="channels_first"):
        super().__init__()
        if dilation != 1 and kernel_size != 3:
            raise RuntimeError("When the dilation isn't 1,the kernel_size should be 3.")
        self.is_vd_mode = is_vd_mode
        self._pool2d_avg = nn.AvgPool2d(kernel_size=2, stride=2, padding="SAME", data_format=data_format)
        self._conv = nn.GroupConv2d(in_channels=in_channels, out_channels=out_channels, kernel_size=kernel_size, stride=stride,
                                    padding=(kernel_size - 1) // 2 if dilation == 1 else dilation, dilation=dilation,
                                    data_format=data_format, b_init=False, n_group=groups)
        self.batch_norm = nn.BatchNorm2d(num_features=out_channels, data_format=data_format)
        if act not in (None, "relu"):
            raise NotImplementedError(act)
        self._act_op = nn.ReLU() if act == "relu" else None

    def forward(self, inputs):
        if self.is_vd_mode:
            inputs = self._pool2d_avg(inputs)
        y = self.batch_norm(self._conv(inputs))
        return self._act_op(y) if self._act_op is not None else y


class BottleneckBlock(nn.Module):
    def __init__(self, in_channels, out_channels, stride, shortcut=True, if_first=False, dilation=1,
                 data_format="channels_first"):
        super().__init__()
        self.conv0 = ConvBNLayer(in_channels, out_channels, 1, act="relu", data_format=data_format)
        self.conv1 = ConvBNLayer(out_channels, out_channels, 3, stride=stride, act="relu", dilation=dilation,
                                 data_format=data_format)
        self.conv2 = ConvBNLayer(out_channels, out_channels * 4, 1, act=None, data_format=data_format)
        if not shortcut:
            self.short = ConvBNLayer(in_channels, out_channels * 4, 1, stride=1,
                                     is_vd_mode=not (if_first or stride == 1), data_format=data_format)
        self.shortcut = shortcut
        self.relu = nn.ReLU()

    def forward(self, inputs):
        y = self.conv2(self.conv1(self.conv0(inputs)))
        short = inputs if self.shortcut else self.short(inputs)
        return self.relu(add(short, y))


class BasicBlock(nn.Module):
    def __init__(self, in_channels, out_channels, stride, dilation=1, shortcut=True, if_first=False,
                 data_format="channels_first"):
        super().__init__()
        self.conv0 = ConvBNLayer(in_channels, out_channels, 3, stride=stride, dilation=dilation, act="relu",
                                 data_format=data_format)
        self.conv1 = ConvBNLayer(out_channels, out_channels, 3, dilation=dilation, act=None, data_format=data_format)
        if not shortcut:
            self.short = ConvBNLayer(in_channels, out_channels, 1, stride=1, is_vd_mode=not (if_first or stride == 1),
                                     data_format=data_format)
        self.shortcut = shortcut
        self.relu = nn.ReLU()

    def forward(self, inputs):
        y = self.conv1(self.conv0(inputs))
        short = inputs if self.shortcut else self.short(inputs)
        return self.relu(add(short, y))


class ResNet_vd(nn.Module):
    def __init__(self, layers=50, output_stride=8, multi_grid=(1, 1, 1), in_channels=3, data_format="channels_first"):
        super().__init__()
        depth = {18: [2, 2, 2, 2], 34: [3, 4, 6, 3], 50: [3, 4, 6, 3], 101: [3, 4, 23, 3], 152: [3, 8, 36, 3],
                 200: [3, 12, 48, 3]}[layers]
        self.layers, self.conv1_logit = layers, None
        num_channels = [64, 256, 512, 1024] if layers >= 50 else [64, 64, 128, 256]
        num_filters = [64, 128, 256, 512]
        self.feat_channels = [c * 4 for c in num_filters] if layers >= 50 else num_filters
        dilation_dict = {8: {2: 2, 3: 4}, 16: {3: 2}}.get(output_stride)
        self.conv1_1 = ConvBNLayer(in_channels, 32, 3, stride=2, act="relu", data_format=data_format)
        self.conv1_2 = ConvBNLayer(32, 32, 3, stride=1, act="relu", data_format=data_format)
        self.conv1_3 = ConvBNLayer(32, 64, 3, stride=1, act="relu", data_format=data_format)
        self.pool2d_max = nn.MaxPool2d(kernel_size=3, stride=2, padding=1, data_format=data_format)
        self.stage_list = []
        for block in range(len(depth)):
            shortcut, block_list = False, []
            for i in range(depth[block]):
                rate = dilation_dict[block] if dilation_dict and block in dilation_dict else 1
                if block == 3:
                    rate = rate * multi_grid[i]
                stride = 2 if i == 0 and block != 0 and rate == 1 else 1
                if layers >= 50:
                    blk = BottleneckBlock(num_channels[block] if i == 0 else num_filters[block] * 4, num_filters[block], stride,
                                          shortcut=shortcut, if_first=block == i == 0, dilation=rate, data_format=data_format)
                else:
                    blk = BasicBlock(num_channels[block] if i == 0 else num_filters[block], num_filters[block], stride,
                                     dilation=rate, shortcut=shortcut, if_first=block == i == 0, data_format=data_format)
                block_list.append(blk)
                shortcut = True
            self.stage_list.append(block_list)

    def forward(self, inputs):
        y = self.pool2d_max(self.conv1_3(self.conv1_2(self.conv1_1(inputs))))
        feat_list = []
        for stage in self.stage_list:
            for block in stage:
                y = block(y)
            feat_list.append(y)
        return feat_list
