"""Functional CPU restatement of the six hot-path topologies.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Plain ``torch.nn.functional`` in
fp32; every function cites the reference file:line it follows.  Parameters are
looked up in a state dict under the reference's own names, so the same seeded
weights drive the reference file (through ``ref_loader``), this restatement and
the B200 model.

Pinned by tests/test_oracle.py against tests/golden/*.npz (logits / feature
digests produced by the reference's own model files, tests/golden/make_golden.py)
and, in the build container, bit-for-bit against ``ref_loader``.
"""
from __future__ import annotations

import torch
import torch.nn.functional as F

EPS = 1e-5  # tensorlayerx BatchNorm epsilon default


class _P:
    """State-dict view with a name prefix."""

    def __init__(self, sd, prefix="", taps=None):
        self.sd, self.prefix, self.taps = sd, prefix, taps

    def sub(self, name):
        return _P(self.sd, f"{self.prefix}{name}.", self.taps)

    def get(self, leaf):
        return self.sd.get(self.prefix + leaf)

    def __getitem__(self, leaf):
        return self.sd[self.prefix + leaf]

    def tap(self, x):
        if self.taps is not None:
            self.taps[self.prefix[:-1]] = x
        return x


def conv(p, x, stride=1, padding=0, groups=1, dilation=1):
    """GroupConv2d -> F.conv2d (cross-correlation, OIHW ``filters``)."""
    return F.conv2d(x, p["filters"], p.get("biases"), stride, padding, dilation, groups)


def bn(p, x, act=None):
    """BatchNorm (eval): (x-mean)/sqrt(var+eps)*gamma+beta, then optional act."""
    y = F.batch_norm(x, p["moving_mean"], p["moving_var"], p["gamma"], p["beta"], False, 0.0, EPS)
    if act == "relu":
        y = F.relu(y)
    return p.tap(y)


def linear(p, x):
    y = torch.matmul(x, p["weights"])
    b = p.get("biases")
    return y if b is None else y + b


# --------------------------------------------------------------------------- #
# ResNet (classification/resnet.py)
# --------------------------------------------------------------------------- #
_RESNET_CFG = {18: [2, 2, 2, 2], 34: [3, 4, 6, 3], 50: [3, 4, 6, 3], 101: [3, 4, 23, 3], 152: [3, 8, 36, 3]}


def _resnet_basic(p, x, stride, has_ds):
    # resnet.py:66-77
    out = F.relu(bn(p.sub("bn1"), conv(p.sub("conv1"), x, stride, 1)))
    out = bn(p.sub("bn2"), conv(p.sub("conv2"), out, 1, 1))
    idt = x
    if has_ds:
        idt = bn(p.sub("downsample.1"), conv(p.sub("downsample.0"), x, stride, 0))
    return p.tap(F.relu(out + idt))


def _resnet_bottleneck(p, x, stride, has_ds, groups):
    # resnet.py:142-156 ; v1.5: the stride sits on the 3x3 (resnet.py:111-121)
    out = F.relu(bn(p.sub("bn1"), conv(p.sub("conv1"), x)))
    out = F.relu(bn(p.sub("bn2"), conv(p.sub("conv2"), out, stride, 1, groups)))
    out = bn(p.sub("bn3"), conv(p.sub("conv3"), out))
    idt = x
    if has_ds:
        idt = bn(p.sub("downsample.1"), conv(p.sub("downsample.0"), x, stride, 0))
    return p.tap(F.relu(out + idt))


def resnet(sd, x, depth=50, groups=1, num_classes=1000, with_pool=True, taps=None):
    """resnet.py:286-300 (ctor :173-237, _make_layer :239-284)."""
    p = _P(sd, "", taps)
    bottleneck = depth >= 50
    expansion = 4 if bottleneck else 1
    x = F.relu(bn(p.sub("bn1"), conv(p.sub("conv1"), x, 2, 3)))        # :287-289
    x = F.max_pool2d(x, 3, 2, 1)                                        # :290 (pads with -inf)
    if taps is not None:
        taps["maxpool"] = x
    in_ch = 64
    for li, (planes, nblk) in enumerate(zip([64, 128, 256, 512], _RESNET_CFG[depth])):
        for bi in range(nblk):
            stride = 2 if (bi == 0 and li > 0) else 1
            has_ds = bi == 0 and (stride != 1 or in_ch != planes * expansion)   # :246
            bp = p.sub(f"layer{li + 1}.{bi}")
            if bottleneck:
                x = _resnet_bottleneck(bp, x, stride, has_ds, groups)
            else:
                x = _resnet_basic(bp, x, stride, has_ds)
            in_ch = planes * expansion
    if with_pool:
        x = F.adaptive_avg_pool2d(x, (1, 1))                            # :295-296
    if num_classes > 0:
        x = linear(p.sub("fc"), x.reshape(x.shape[0], -1))              # :298-299
    return x


# --------------------------------------------------------------------------- #
# ResNeXt (classification/resnext.py)
# --------------------------------------------------------------------------- #
def _resnext_cbl(p, x, k, stride=1, groups=1, act=None):
    # resnext.py:54-57 ; padding=(k-1)//2 (:36)
    return bn(p.sub("batch_norm"), conv(p.sub("_conv"), x, stride, (k - 1) // 2, groups), act)


def resnext(sd, x, layers=50, cardinality=32, taps=None):
    """resnext.py:201-209 (ctor :123-199, block :109-119)."""
    p = _P(sd, "", taps)
    depth = {50: [3, 4, 6, 3], 101: [3, 4, 23, 3], 152: [3, 8, 36, 3]}[layers]
    y = _resnext_cbl(p.sub("conv"), x, 7, 2, act="relu")                # :150-158
    y = F.max_pool2d(y, 3, 2, 1)                                        # :159-164
    for block in range(4):
        for i in range(depth[block]):
            bp = p.sub(f"bb_{block}_{i}")
            stride = 2 if (i == 0 and block != 0) else 1                # :183
            t = _resnext_cbl(bp.sub("conv0"), y, 1, act="relu")
            t = _resnext_cbl(bp.sub("conv1"), t, 3, stride, cardinality, act="relu")
            t = _resnext_cbl(bp.sub("conv2"), t, 1)
            short = y if i != 0 else _resnext_cbl(bp.sub("short"), y, 1, stride)   # :113-116
            y = bp.tap(F.relu(short + t))                               # :117-118
    y = F.adaptive_avg_pool2d(y, 1)
    y = y.reshape(-1, 2048)                                             # :207
    return linear(p.sub("out"), y)


# --------------------------------------------------------------------------- #
# MobileNetV2 (classification/mobilenetv2.py + ops/ops_fusion.py + utils/common_func.py)
# --------------------------------------------------------------------------- #
def _make_divisible(v, divisor=8):
    # utils/common_func.py:1-16
    new_v = max(divisor, int(v + divisor / 2) // divisor * divisor)
    if new_v < 0.9 * v:
        new_v += divisor
    return new_v


def _cna(p, x, k=3, stride=1, groups=1, act="relu6"):
    # ops_fusion.py:39-48 : Sequential(conv(pad=(k-1)//2), bn, act)
    y = bn(p.sub("1"), conv(p.sub("0"), x, stride, (k - 1) // 2, groups))
    if act == "relu6":
        y = F.relu6(y)
    elif act == "relu":
        y = F.relu(y)
    return y


def mobilenet_v2(sd, x, scale=1.0, taps=None):
    """mobilenetv2.py:102-109 (ctor :67-100, block :15-40)."""
    p = _P(sd, "", taps)
    setting = [[1, 16, 1, 1], [6, 24, 2, 2], [6, 32, 3, 2], [6, 64, 4, 2], [6, 96, 3, 1], [6, 160, 3, 2],
               [6, 320, 1, 1]]                                          # :76-78
    in_ch = _make_divisible(32 * scale)
    x = _cna(p.sub("features.0"), x, 3, 2)                              # :82-83
    idx = 1
    for t, c, n, s in setting:
        out_ch = _make_divisible(c * scale)
        for i in range(n):
            stride = s if i == 0 else 1
            hidden = int(round(in_ch * t))
            bp = p.sub(f"features.{idx}.conv")
            y, j = x, 0
            if t != 1:                                                  # :24-26
                y = _cna(bp.sub("0"), y, 1)
                j = 1
            y = _cna(bp.sub(str(j)), y, 3, stride, hidden)              # depthwise :27-29
            y = bn(bp.sub(str(j + 2)), conv(bp.sub(str(j + 1)), y))     # linear project :29-33
            x = (x + y) if (stride == 1 and in_ch == out_ch) else y     # :37-40
            bp.tap(x)
            in_ch = out_ch
            idx += 1
    x = _cna(p.sub(f"features.{idx}"), x, 1)                            # :91-92
    x = F.adaptive_avg_pool2d(x, 1)
    x = torch.flatten(x, 1)                                             # :107
    return linear(p.sub("classifier.1"), x)                             # Dropout(0.2) is identity in eval


# --------------------------------------------------------------------------- #
# MobileNetV1 (classification/mobilenetv1.py)
# --------------------------------------------------------------------------- #
def mobilenet_v1(sd, x, scale=1.0, taps=None):
    """mobilenetv1.py:254-262 (ctor :116-252, DepthwiseSeparable :68-102)."""
    p = _P(sd, "", taps)
    cfg = [(32, 64, 1), (64, 128, 2), (128, 128, 1), (128, 256, 2), (256, 256, 1), (256, 512, 2)] + \
          [(512, 512, 1)] * 5 + [(512, 1024, 2), (1024, 1024, 1)]
    x = _cna(p.sub("conv1"), x, 3, 2, act="relu")                       # :125-132
    for i, (c1, c2, s) in enumerate(cfg):
        bp = p.sub(f"dwsl.{i}")
        x = _cna(bp.sub("_depthwise_conv"), x, 3, s, int(c1 * scale), act="relu")
        x = _cna(bp.sub("_pointwise_conv"), x, 1, act="relu")
    x = F.adaptive_avg_pool2d(x, 1)
    x = x.reshape(x.shape[0], -1)                                       # :260
    return linear(p.sub("fc"), x)


# --------------------------------------------------------------------------- #
# DarkNet-53 classifier (classification/darknet53.py) — ReLU
# --------------------------------------------------------------------------- #
def _dk_cls_cbl(p, x, k, stride, pad):
    # darknet53.py:35-38 ; BatchNorm(act='relu') :30-33
    return bn(p.sub("_bn"), conv(p.sub("_conv"), x, stride, pad), "relu")


def darknet53_cls(sd, x, taps=None):
    """darknet53.py:100-133."""
    p = _P(sd, "", taps)
    x = _dk_cls_cbl(p.sub("_conv1"), x, 3, 1, 1)
    x = _dk_cls_cbl(p.sub("_conv2"), x, 3, 2, 1)
    for si, n in enumerate([1, 2, 8, 8, 4]):                            # :60
        for bi in range(1, n + 1):
            bp = p.sub(f"_basic_block_{si}{bi}")
            y = _dk_cls_cbl(bp.sub("_conv1"), x, 1, 1, 0)               # :50-53
            y = _dk_cls_cbl(bp.sub("_conv2"), y, 3, 1, 1)
            x = bp.tap(x + y)
        if si < 4:
            x = _dk_cls_cbl(p.sub(f"_downsample_{si}"), x, 3, 2, 1)
    x = F.adaptive_avg_pool2d(x, 1)
    x = x.squeeze(3).squeeze(2)                                         # :131
    return linear(p.sub("_out"), x)


# --------------------------------------------------------------------------- #
# DarkNet-53 detection backbone (detection/backbones/darknet.py) — LeakyReLU(0.1)
# --------------------------------------------------------------------------- #
def _dk_det_cbl(p, x, stride, pad):
    # darknet.py:54-58
    y = bn(p.sub("batch_norm"), conv(p.sub("conv"), x, stride, pad))
    return F.leaky_relu(y, 0.1)


def darknet53_det(sd, inputs, return_idx=(2, 3, 4), taps=None):
    """darknet.py:299-312 ; input is a dict {"images": NCHW} (:300); returns a list."""
    p = _P(sd, "", taps)
    x = inputs["images"] if isinstance(inputs, dict) else inputs
    out = _dk_det_cbl(p.sub("conv0"), x, 1, 1)
    out = _dk_det_cbl(p.sub("downsample0.conv_bn_layer"), out, 2, 1)
    blocks = []
    for i, n in enumerate([1, 2, 8, 8, 4]):                             # :217
        sp = p.sub(f"darknet_conv_block_list.{i}")
        for bi in range(n):
            bp = sp.sub("basicblock0") if bi == 0 else sp.sub(f"res_blocks.{bi - 1}")
            y = _dk_det_cbl(bp.sub("conv1"), out, 1, 0)                 # :155-159
            y = _dk_det_cbl(bp.sub("conv2"), y, 1, 1)
            out = bp.tap(out + y)
        if i in return_idx:
            blocks.append(out)
        if i < 4:
            out = _dk_det_cbl(p.sub(f"downsample_list.{i}.conv_bn_layer"), out, 2, 1)
    return blocks


def yolov3_darknet53(sd, inputs, taps=None):
    """DarkNet-53 -> YOLOv3FPN -> YOLOv3Head output convs (detection/yolov3.py:51-68 without post-processing).

    YoloDetBlock :178-180 (conv_module = 1x1, 3x3, 1x1, 3x3, 1x1 ConvBN+LeakyReLU; tip = 3x3), YOLOv3FPN.forward
    :233-258 (deepest map first; route -> 1x1 ConvBN -> nearest x2 -> concat([route, x], axis=1)), YOLOv3Head.forward
    :353 (1x1 conv with bias per level)."""
    p = _P(sd, "", taps)
    body = darknet53_det(_SubDict(sd, "backbone."), inputs)
    neck, route = [], None
    for i, x in enumerate(body[::-1]):
        if i > 0:
            x = torch.cat([route, x], dim=1)                            # :244
        bp = p.sub(f"neck.yolo_blocks.{i}")
        for j, k in enumerate([1, 3, 1, 3, 1]):                         # :146-152
            x = _dk_det_cbl(bp.sub(f"conv_module.{j}"), x, 1, (k - 1) // 2)
        route = x
        neck.append(_dk_det_cbl(bp.sub("tip"), route, 1, 1))            # :179
        if i < 2:
            route = _dk_det_cbl(p.sub(f"neck.routes.{i}"), route, 1, 0)     # :251
            route = F.interpolate(route, scale_factor=2.0)              # :252-253 (nearest)
    heads = [conv(p.sub(f"yolo_head.yolo_outputs.{i}"), f) for i, f in enumerate(neck)]    # :353
    return {"body_feats": body, "neck_feats": neck, "yolo_head_outs": heads}


class _SubDict:
    """Read-only view of a state dict under a key prefix."""

    def __init__(self, sd, prefix):
        self.sd, self.prefix = sd, prefix

    def __getitem__(self, k):
        return self.sd[self.prefix + k]

    def get(self, k, default=None):
        return self.sd.get(self.prefix + k, default)


def mobilenet_v1_det(sd, inputs, feature_maps=(4, 6, 13), taps=None):
    """detection/backbones/mobilenet_v1.py:238-245 (ConvBNLayer :44-50: conv, BN, ReLU; DepthwiseSeparable :99-103)."""
    p = _P(sd, "", taps)
    x = inputs["images"] if isinstance(inputs, dict) else inputs

    def cbl(q, x, k, stride, groups=1):
        return F.relu(bn(q.sub("my_batch_norm"), conv(q.sub("_conv"), x, stride, (k - 1) // 2, groups)))

    cfg = [(32, 64, 1), (64, 128, 2), (128, 128, 1), (128, 256, 2), (256, 256, 1), (256, 512, 2)] + \
          [(512, 512, 1)] * 5 + [(512, 1024, 2), (1024, 1024, 1)]        # :199-209
    y = cbl(p.sub("conv1"), x, 3, 2)
    outs = []
    for idx, (c1, c2, s) in enumerate(cfg, start=1):
        bp = p.sub(f"dwsl.{idx - 1}")
        y = cbl(bp.sub("_depthwise_conv"), y, 3, s, c1)
        y = cbl(bp.sub("_pointwise_conv"), y, 1, 1)
        if idx in feature_maps:
            outs.append(y)
    return outs


def resnet_vd(sd, x, layers=50, output_stride=8, taps=None):
    """segmentation/backbones/resnet_vd.py:315-326 (ConvBNLayer :44-50, BottleneckBlock :103-113, BasicBlock :160-169,
    stage table :222-311)."""
    p = _P(sd, "", taps)
    depth = {18: [2, 2, 2, 2], 34: [3, 4, 6, 3], 50: [3, 4, 6, 3], 101: [3, 4, 23, 3]}[layers]
    dil = {8: {2: 2, 3: 4}, 16: {3: 2}}.get(output_stride, {})

    def cbl(q, x, k, stride=1, dilation=1, vd=False, act=None):
        if vd:
            x = F.avg_pool2d(x, 2, 2)                                   # :45-46
        y = bn(q.sub("batch_norm"), conv(q.sub("_conv"), x, stride, (k - 1) // 2 if dilation == 1 else dilation, 1, dilation))
        return F.relu(y) if act == "relu" else y

    y = cbl(p.sub("conv1_1"), x, 3, 2, act="relu")
    y = cbl(p.sub("conv1_2"), y, 3, act="relu")
    y = cbl(p.sub("conv1_3"), y, 3, act="relu")
    y = F.max_pool2d(y, 3, 2, 1)
    feats = []
    for block in range(4):
        for i in range(depth[block]):
            rate = dil.get(block, 1)
            stride = 2 if i == 0 and block != 0 and rate == 1 else 1
            bp = p.sub(f"stage_list.{block}.{i}")
            vd = i == 0 and not (block == 0 or stride == 1)             # short: is_vd_mode = not (if_first or stride == 1)
            if layers >= 50:
                t = cbl(bp.sub("conv0"), y, 1, act="relu")
                t = cbl(bp.sub("conv1"), t, 3, stride, rate, act="relu")
                t = cbl(bp.sub("conv2"), t, 1)
            else:
                t = cbl(bp.sub("conv0"), y, 3, stride, rate, act="relu")
                t = cbl(bp.sub("conv1"), t, 3, 1, rate)
            short = y if i != 0 else cbl(bp.sub("short"), y, 1, 1, 1, vd)
            y = bp.tap(F.relu(short + t))
        feats.append(y)
    return feats


def resnest(sd, x, layers=(3, 4, 6, 3), radix=2, cardinality=1, taps=None):
    """classification/resnest.py:672-682 with the resnest50 / resnest101 options (:707-733: radix 2, deep stem, avg_down,
    avd after the SplatConv): ConvBNLayer :47-50, rSoftmax :64-82, SplatConv :146-166, BottleneckBlock :312-328."""
    p = _P(sd, "", taps)

    def cbl(q, x, k, stride=1, groups=1, act=None):
        return bn(q.sub("batch_norm"), conv(q.sub("_conv"), x, stride, (k - 1) // 2, groups), act)

    def splat(q, x):
        y = cbl(q.sub("conv1"), x, 3, 1, cardinality * radix, "relu")            # :147
        parts = torch.chunk(y, radix, dim=1)                                     # :149-150
        gap = parts[0]
        for t in parts[1:]:
            gap = gap + t                                                        # add_n :151
        gap = F.adaptive_avg_pool2d(gap, 1)                                      # :154
        att = conv(q.sub("conv3"), cbl(q.sub("conv2"), gap, 1, 1, cardinality, "relu"), 1, 0, cardinality)   # :155-156
        b = att.shape[0]
        att = att.reshape(b, cardinality, radix, -1).permute(0, 2, 1, 3)         # rSoftmax :72-75
        att = F.softmax(att, dim=1).reshape(b, -1, 1, 1)                         # :76-77
        out = None
        for a, t in zip(torch.chunk(att, radix, dim=1), parts):                  # :159-162
            out = t * a if out is None else out + t * a
        return out

    y = cbl(p.sub("stem.conv1"), x, 3, 2, 1, "relu")
    y = cbl(p.sub("stem.conv2"), y, 3, 1, 1, "relu")
    y = cbl(p.sub("stem.conv3"), y, 3, 1, 1, "relu")
    y = F.max_pool2d(y, 3, 2, 1)
    inplanes = y.shape[1]
    for li, n in enumerate(layers):
        planes = 64 << li
        for i in range(n):
            stride = 2 if (i == 0 and li > 0) else 1
            is_first = i == 0 and li > 0                                         # ResNeStLayer(is_first=...) :546,560 (default True)
            bp = p.sub(f"layer{li + 1}.layer{li + 1}_bottleneck_{i}")
            t = cbl(bp.sub("conv1"), y, 1, 1, 1, "relu")
            t = splat(bp.sub("conv2"), t)
            if stride > 1 or is_first:
                t = F.avg_pool2d(t, 3, stride, 1)                                # avd, after the SplatConv (:318-319)
            t = cbl(bp.sub("conv3"), t, 1)
            short = y
            if stride != 1 or inplanes != planes * 4:
                short = F.avg_pool2d(short, stride, stride, 0)                   # avg_down :321-322
                short = bn(bp.sub("batch_norm"), conv(bp.sub("conv4"), short))
            y = bp.tap(F.relu(short + t))
            inplanes = planes * 4
    y = torch.flatten(F.adaptive_avg_pool2d(y, 1), 1)
    return linear(p.sub("out"), y)


FORWARD = {
    "resnest50": lambda sd, x, **k: resnest(sd, x, (3, 4, 6, 3), **k),
    "resnest101": lambda sd, x, **k: resnest(sd, x, (3, 4, 23, 3), **k),
    "resnet18": lambda sd, x, **k: resnet(sd, x, 18, **k),
    "resnet34": lambda sd, x, **k: resnet(sd, x, 34, **k),
    "resnet50": lambda sd, x, **k: resnet(sd, x, 50, **k),
    "resnet101": lambda sd, x, **k: resnet(sd, x, 101, **k),
    "wide_resnet50_2": lambda sd, x, **k: resnet(sd, x, 50, **k),
    "resnext50_32x4d": lambda sd, x, **k: resnext(sd, x, 50, 32, **k),
    "resnext50_64x4d": lambda sd, x, **k: resnext(sd, x, 50, 64, **k),
    "mobilenet_v1": mobilenet_v1,
    "mobilenet_v2": mobilenet_v2,
    "darknet53_cls": darknet53_cls,
    "darknet53_det": darknet53_det,
    "yolov3_darknet53": yolov3_darknet53,
    "mobilenet_v1_det": mobilenet_v1_det,
    "resnet50_vd": lambda sd, x, **k: resnet_vd(sd, x, 50, **k),
    "resnet18_vd": lambda sd, x, **k: resnet_vd(sd, x, 18, **k),
}


def forward(name, sd, x, **kw):
    """Run restated model ``name`` in eval mode under ``torch.no_grad``."""
    with torch.no_grad():
        return FORWARD[name](sd, x, **kw)
