"""YOLOv3 neck and head convolutions on the B200 path.

Mirrors tlxcv/models/detection/yolov3.py: ``YoloDetBlock`` (:122-180, five 1x1 / 3x3 ConvBN + LeakyReLU layers and the
3x3 ``tip``), ``YOLOv3FPN`` (:183-258: deepest map first; ``route`` = 1x1 transition conv, nearest x2 up-sampling,
``tlx.concat([route, x], axis=1)`` with the next backbone map) and the output convolutions of ``YOLOv3Head``
(:261-378: 1x1, **bias**, no BatchNorm, ``len(anchors) * (num_classes + 5)`` = 291 channels).  Module paths follow the
reference's attribute names; the list-held sub-modules (``yolo_blocks``, ``routes``, ``yolo_outputs`` are plain Python
lists in the reference) register as ``<attr>.<i>`` like the DarkNet stages do.

In a plan the up-sampling and the concat are ONE memory pass (``TLXCV_OP_UPSAMPLE_CONCAT``), every ConvBN + LeakyReLU
is one tcgen05 launch, and the head maps leave as NCHW fp32.  Box decoding and NMS (``BBoxPostProcess``, host-side numpy
in the reference) and the training loss are outside the path: ``YOLOv3.forward`` returns the raw head outputs.
"""
from __future__ import annotations

from .. import concat, nn
from .darknet import ConvBNLayer, DarkNet

__all__ = ["YoloDetBlock", "YOLOv3FPN", "YOLOv3Head", "YOLOv3"]


class Interpolater:
    """detection/utils/layers.py:132-139: nearest up-sampling of the traced map."""

    def __init__(self, data_format="channels_first"):
        self.data_format = data_format

    def __call__(self, x, scale_factor=2.0, **kwds):
        return x.graph.interpolate(x, scale_factor=scale_factor, mode=kwds.get("mode", "nearest"))


class YoloDetBlock(nn.Module):
    def __init__(self, ch_in, channel, norm_type="bn", freeze_norm=False, name="", data_format="channels_first"):
        super().__init__()
        if channel % 2:
            raise ValueError(f"channel {channel} cannot be divided by 2")
        self.ch_in, self.channel = ch_in, channel
        conv_def = [(ch_in, channel, 1), (channel, channel * 2, 3), (channel * 2, channel, 1), (channel, channel * 2, 3),
                    (channel * 2, channel, 1)]                                  # conv0, conv1, conv2, conv3, route (:146-152)
        self.conv_module = nn.Sequential([ConvBNLayer(ci, co, k, padding=(k - 1) // 2, data_format=data_format)
                                          for ci, co, k in conv_def])
        self.tip = ConvBNLayer(channel, channel * 2, 3, padding=1, data_format=data_format)

    def forward(self, inputs):
        route = self.conv_module(inputs)
        return route, self.tip(route)


class YOLOv3FPN(nn.Module):
    def __init__(self, in_channels=(256, 512, 1024), norm_type="bn", freeze_norm=False, data_format="channels_first"):
        super().__init__()
        in_channels = list(in_channels)
        if not in_channels:
            raise ValueError("in_channels length should > 0")
        self.interpolate = Interpolater(data_format)
        self.in_channels, self.num_blocks, self._out_channels = in_channels, len(in_channels), []
        self.yolo_blocks, self.routes = [], []
        for i, in_channel in enumerate(in_channels[::-1]):
            if i > 0:
                in_channel += 512 // 2 ** i                                     # the up-sampled route joins (:218-219)
            self.yolo_blocks.append(YoloDetBlock(in_channel, channel=512 // 2 ** i, data_format=data_format))
            self._out_channels.append(1024 // 2 ** i)
            if i < self.num_blocks - 1:
                self.routes.append(ConvBNLayer(512 // 2 ** i, 256 // 2 ** i, 1, 1, padding=0, data_format=data_format))

    def forward(self, X, for_mot=False):
        if for_mot:
            raise NotImplementedError("for_mot (embedding features) is not on the B200 path")
        if len(X) != self.num_blocks:
            raise ValueError(f"YOLOv3FPN expects {self.num_blocks} maps")
        X = X[::-1]
        yolo_feats = []
        route = None
        for i, x in enumerate(X):
            if i > 0:
                x = concat([route, x], axis=1)
            route, tip = self.yolo_blocks[i](x)
            yolo_feats.append(tip)
            if i < self.num_blocks - 1:
                route = self.interpolate(self.routes[i](route), scale_factor=2.0)
        return yolo_feats


_ANCHORS = [[10, 13], [16, 30], [33, 23], [30, 61], [62, 45], [59, 119], [116, 90], [156, 198], [373, 326]]


class YOLOv3Head(nn.Module):
    def __init__(self, in_channels=(1024, 512, 256), anchors=_ANCHORS, anchor_masks=((6, 7, 8), (3, 4, 5), (0, 1, 2)),
                 num_classes=92, loss=None, batch_transforms=None, iou_aware=False, iou_aware_factor=0.4,
                 data_format="channels_first"):
        super().__init__()
        if iou_aware:
            raise NotImplementedError("iou_aware heads are not on the B200 path")
        self.in_channels, self.num_classes = list(in_channels), num_classes
        self.anchors = [[list(anchors[i]) for i in mask] for mask in anchor_masks]          # parse_anchor (:335-343)
        self.mask_anchors = [[v for i in mask for v in anchors[i]] for mask in anchor_masks]
        self.num_outputs = len(self.anchors)
        self.yolo_outputs = []
        for i, a in enumerate(self.anchors):
            self.yolo_outputs.append(nn.GroupConv2d(in_channels=self.in_channels[i], out_channels=len(a) * (num_classes + 5),
                                                    kernel_size=1, stride=1, padding=0, data_format=data_format,
                                                    b_init=nn.initializers.xavier_uniform()))

    def forward(self, outputs, targets=None):
        if targets is not None:
            raise NotImplementedError("the YOLOv3 training loss is outside the B200 inference path")
        feats = outputs["neck_feats"] if isinstance(outputs, dict) else outputs
        if len(feats) != len(self.anchors):
            raise ValueError(f"YOLOv3Head expects {len(self.anchors)} maps")
        return [fn(feat) for fn, feat in zip(self.yolo_outputs, feats)]


class YOLOv3(nn.Module):
    """DarkNet-53 -> YOLOv3FPN -> YOLOv3Head (detection/yolov3.py:22-104), up to the raw head maps:
    ``{"body_feats", "neck_feats", "yolo_head_outs"}``.  Decoding + NMS run on the host in the reference and are not built."""

    def __init__(self, backbone="DarkNet", data_format="channels_first", for_mot=False, num_classes=92):
        super().__init__()
        if for_mot:
            raise NotImplementedError("for_mot is not on the B200 path")
        self.backbone = DarkNet(data_format=data_format) if isinstance(backbone, str) else backbone
        self.neck = YOLOv3FPN(data_format=data_format)
        self.yolo_head = YOLOv3Head(num_classes=num_classes, data_format=data_format)
        self.data_format = data_format

    def forward(self, inputs):
        body_feats = self.backbone(inputs)
        neck_feats = self.neck(body_feats)
        return {"body_feats": body_feats, "neck_feats": neck_feats, "yolo_head_outs": self.yolo_head(neck_feats)}
