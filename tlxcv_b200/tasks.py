"""``tlxcv.tasks.ImageClassification`` on the B200 path (tasks/image_classification.py:6-23).

``predict`` is ``set_eval()`` + backbone + ``argmax(axis=-1)``; the argmax is part
of the same plan, so only ``[N]`` int64 leaves the device.  ``loss_fn`` belongs
to training and is outside this path.
"""
from __future__ import annotations

from . import argmax, losses, nn


class ImageClassification(nn.Module):
    def __init__(self, backbone):
        super().__init__()
        self.backbone = backbone
        self._predictor = _Predict(backbone)
        self._loss = _Loss()

    def loss_fn(self, output, target):
        """``tlx.losses.softmax_cross_entropy_with_logits(output, target)`` (tasks/image_classification.py:10-15): the mean
        cross-entropy of fp32 logits against int64 labels, evaluated on the device (validation loss; there is no backward)."""
        if hasattr(output, "graph"):                 # called inside a traced forward: part of that plan
            return losses.softmax_cross_entropy_with_logits(output, target)
        return self._loss(output, target)

    def forward(self, inputs):
        return self.backbone(inputs)

    def predict(self, inputs):
        self.set_eval()
        return self._predictor(inputs)

    def state_dict(self, *args, **kwargs):
        sd = super().state_dict(*args, **kwargs)
        return type(sd)((k, v) for k, v in sd.items() if "_predictor." not in k and "_loss." not in k)


class _Predict(nn.Module):
    """backbone + argmax traced as one plan."""

    def __init__(self, backbone):
        super().__init__()
        object.__setattr__(self, "_bb", backbone)      # not registered: the backbone belongs to the task module

    def set_eval(self):
        self.is_train = False
        return self

    def named_modules(self, *a, **k):
        yield from self._bb.named_modules(*a, **k)

    def forward(self, inputs):
        return argmax(self._bb(inputs), axis=-1)


class _Loss(nn.Module):
    def forward(self, output, target):
        return losses.softmax_cross_entropy_with_logits(output, target)
