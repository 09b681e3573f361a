"""Device-side image pre-processing on the hot path.

The reference's demos prepare every image on the host with ``tensorlayerx.vision.transforms``:
``Compose([Resize, Normalize(mean, std), ToTensor(data_format)])`` followed by ``expand_dims``
(demo/image_classification/predict-resnet.py:50-56) — uint8 HWC pixels become a normalised fp32 CHW tensor
that is four times larger than the image and has to cross PCIe.  ``NormalizeToTensor`` is the batched,
fused equivalent: the forward takes the (already resized) uint8 ``(N, H, W, C)`` batch on the device and the
``(x - mean) / std`` + layout change happens inside the plan's input kernel (``TLXCV_OP_IMPORT_U8_NHWC``),
writing the bf16 NHWC activation the stem conv reads.  Put it in front of any backbone:

    net = vision.Preprocessed(models.resnet50(), mean=(125.31, 122.95, 113.86), std=(62.99, 62.09, 66.70))
    logits = net(uint8_batch_nhwc.cuda())
"""
from __future__ import annotations

import torch

from . import graph as _g
from . import nn


class NormalizeToTensor(nn.Module):
    """``Normalize(mean, std)`` + ``ToTensor('CHW')`` of the reference's transform pipeline, for a uint8 NHWC batch.

    ``mean`` / ``std`` are per-channel, in pixel units (the reference passes e.g. mean=(125.31, 122.95, 113.86))."""

    def __init__(self, mean, std, name=None):
        super().__init__(name=name)
        mean, std = [float(v) for v in mean], [float(v) for v in std]
        if len(mean) != len(std) or not 1 <= len(mean) <= 4:
            raise ValueError("mean and std need one entry per channel (1 to 4 channels)")
        if any(s == 0.0 for s in std):
            raise ValueError("std entries must be non-zero")
        self.register_buffer("mean", torch.tensor(mean, dtype=torch.float32))
        self.register_buffer("std", torch.tensor(std, dtype=torch.float32))

    def forward(self, x):
        g = _g.active()
        if g is None or not isinstance(x, _g.SymTensor):
            raise RuntimeError("NormalizeToTensor only executes inside a traced plan (call the enclosing module with a "
                               "uint8 (N, H, W, C) CUDA tensor)")
        return g.normalize_u8(x, self)


class Resize(nn.Module):
    """``Resize((H, W))`` of the reference's transform pipeline (demo/image_classification/predict-resnet.py:51) for a uint8
    NHWC batch: bilinear, bit-identical to ``cv2.resize(img, (W, H), interpolation=cv2.INTER_LINEAR)`` - what tensorlayerx's
    ``Resize`` runs on a numpy image (oracle/cv_resize.py restates it).  Executes inside the input pass of the
    ``NormalizeToTensor`` that must follow it: the resized uint8 image never exists in memory."""

    def __init__(self, size, interpolation="bilinear", name=None):
        super().__init__(name=name)
        if interpolation != "bilinear":
            raise NotImplementedError("Resize: only bilinear interpolation is on the B200 path")
        self.size = (int(size), int(size)) if isinstance(size, int) else (int(size[0]), int(size[1]))

    def forward(self, x):
        g = _g.active()
        if g is None or not isinstance(x, _g.SymTensor):
            raise RuntimeError("Resize only executes inside a traced plan (call the enclosing module with a uint8 (N, H, W, C) "
                               "CUDA tensor)")
        return g.resize_u8(x, self.size)


class Preprocessed(nn.Module):
    """``backbone(NormalizeToTensor(mean, std)([Resize(size)](images_uint8_nhwc)))`` as one module / one plan."""

    def __init__(self, backbone, mean, std, resize=None, name=None):
        super().__init__(name=name)
        self.resize = Resize(resize) if resize is not None else None
        self.preprocess = NormalizeToTensor(mean, std)
        self.backbone = backbone

    def forward(self, images):
        if self.resize is not None:
            images = self.resize(images)
        return self.backbone(self.preprocess(images))
