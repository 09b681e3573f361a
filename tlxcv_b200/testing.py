"""Seeded synthetic weights and inputs shared by the tests, the bench and smoke().

The reference ships no usable weight recipe: tensorlayerx's default inits
give un-damped residual streams (logit std ~1e3) and, in resnext.py:49-50 /
darknet53.py:30-32, a *xavier-uniform moving variance* that can be negative
(NaN in eval mode).  SURVEY.md App. D defines the recipe implemented here; it
keeps every activation O(1) so that the north-star's absolute logit tolerance
(1e-2 in bf16) is meaningful.

All tensors come from one ``torch.Generator`` in state-dict order, so any two
modules with the same state-dict manifest (the reference file under the oracle
stand-in, the oracle restatement, the B200 model) receive identical weights.
"""
from __future__ import annotations

import re
from collections import OrderedDict

import torch

WEIGHT_SEED = 1234
INPUT_SEED = 0

# last BatchNorm of each residual branch (gamma x 0.25 keeps the residual stream O(1))
_DAMP = {
    "resnet_basic": [r"\.bn2\.gamma$"],
    "resnet_bottleneck": [r"\.bn3\.gamma$"],
    "resnext": [r"\.conv2\.batch_norm\.gamma$"],
    "mobilenet_v1": [],
    # project BN of every inverted-residual block: conv.2 (t=1 block) / conv.3
    "mobilenet_v2": [r"features\.1\.conv\.2\.gamma$", r"features\.\d+\.conv\.3\.gamma$"],
    "darknet53_cls": [r"_basic_block_\d+\._conv2\._bn\.gamma$"],
    "darknet53_det": [r"\.conv2\.batch_norm\.gamma$"],
}

# model name -> (damp family, FC gain chosen so that logits std is ~0.25-0.3 on randn images)
RECIPES = {
    "resnet18": ("resnet_basic", 0.085),
    "resnet34": ("resnet_basic", 0.085),
    "resnet50": ("resnet_bottleneck", 0.07),
    "resnet101": ("resnet_bottleneck", 0.07),
    "resnet152": ("resnet_bottleneck", 0.07),
    "wide_resnet50_2": ("resnet_bottleneck", 0.07),
    "wide_resnet101_2": ("resnet_bottleneck", 0.07),
    "resnext50_32x4d": ("resnext", 0.06),
    "resnext50_64x4d": ("resnext", 0.06),
    "resnext101_32x4d": ("resnext", 0.06),
    "mobilenet_v1": ("mobilenet_v1", 0.11),
    "mobilenet_v2": ("mobilenet_v2", 3.0),
    "darknet53_cls": ("darknet53_cls", 0.02),
    "darknet53_det": ("darknet53_det", 1.0),
}


def seeded_state_dict(manifest, model_name: str, seed: int = WEIGHT_SEED, fc_gain: float | None = None):
    """Build an ``OrderedDict`` name -> fp32 CPU tensor for ``manifest``.

    ``manifest`` is an ordered mapping name -> shape (e.g. ``module.state_dict()``;
    only names and shapes are read).  Leaf names follow tensorlayerx:
    ``filters`` (OIHW), ``biases``, ``gamma``, ``beta``, ``moving_mean``,
    ``moving_var``, ``weights`` (Linear, (in, out)).
    """
    family, gain = RECIPES[model_name]
    if fc_gain is not None:
        gain = fc_gain
    damp = [re.compile(p) for p in _DAMP[family]]
    g = torch.Generator().manual_seed(seed)
    out = OrderedDict()
    for name, ref in manifest.items():
        shape = tuple(ref.shape) if hasattr(ref, "shape") else tuple(ref)
        leaf = name.rsplit(".", 1)[-1]
        if leaf == "filters":
            fan_in = shape[1] * shape[2] * shape[3]
            t = torch.randn(shape, generator=g) * (2.0 / fan_in) ** 0.5
        elif leaf == "gamma":
            t = 0.75 + 0.5 * torch.rand(shape, generator=g)
            if any(p.search(name) for p in damp):
                t = t * 0.25
        elif leaf in ("beta", "biases", "moving_mean"):
            t = torch.randn(shape, generator=g) * 0.05
        elif leaf == "moving_var":
            t = 0.75 + 0.5 * torch.rand(shape, generator=g)
        elif leaf == "weights":
            t = torch.randn(shape, generator=g) * (1.0 / shape[0]) ** 0.5 * gain
        else:
            raise KeyError(f"unknown parameter leaf {name!r}")
        out[name] = t.contiguous()
    return out


def synthetic_images(n: int, size: int = 224, seed: int = INPUT_SEED, channels: int = 3):
    """``torch.randn(N,3,H,W)`` fp32 NCHW: normalised-image statistics
    (cf. demo/image_classification/predict-resnet.py:50-54)."""
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, channels, size, size, generator=g)


def state_dict_digest(sd) -> str:
    """Order-sensitive SHA-256 over names, shapes and raw fp32 bytes."""
    import hashlib

    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(str(tuple(v.shape)).encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()
