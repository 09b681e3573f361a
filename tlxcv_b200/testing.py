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
    "resnet_vd_bottleneck": [r"\.conv2\.batch_norm\.gamma$"],
    "resnest": [r"_bottleneck_\d+\.conv3\.batch_norm\.gamma$"],
    "resnet_vd_basic": [r"stage_list\.\d+\.\d+\.conv1\.batch_norm\.gamma$"],
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
    "yolov3_darknet53": ("darknet53_det", 1.0),
    "mobilenet_v1_det": ("mobilenet_v1", 1.0),
    "resnest50": ("resnest", 0.07),
    "resnest101": ("resnest", 0.07),
    "resnet50_vd": ("resnet_vd_bottleneck", 1.0),
    "resnet18_vd": ("resnet_vd_basic", 1.0),
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


def structured_images(n: int, size: int = 224, seed: int = INPUT_SEED, first: int = 0):
    """Synthetic images that DIFFER from one another the way photographs do: per-image contrast, per-channel brightness
    and a low-frequency pattern on top of the pixel noise.  With ``torch.randn`` alone every image has the same
    statistics, a random-init network pools them to nearly identical features (logit std over images 0.009 against 0.29
    over classes for ResNet-50) and "identical top-1" is satisfied by two classes.  Image ``i`` depends only on
    ``(seed, i)``, so the first images of a large batch equal a small batch (``first`` offsets ``i``)."""
    import math

    yy = torch.linspace(0, 1, size).view(1, size, 1)
    xx = torch.linspace(0, 1, size).view(1, 1, size)
    out = torch.empty(n, 3, size, size)
    for i in range(n):
        g = torch.Generator().manual_seed(seed * 1000003 + first + i)
        par = torch.rand(12, generator=g)
        gain = 0.25 + 1.5 * par[0]
        off = (par[1:4] - 0.5) * 2.4
        fy, fx = 0.5 + 6.0 * par[4], 0.5 + 6.0 * par[5]
        ph = par[6:9] * (2 * math.pi)
        amp = par[9:12] * 1.5
        wave = amp.view(3, 1, 1) * torch.sin(2 * math.pi * (fy * yy + fx * xx) + ph.view(3, 1, 1))
        out[i] = torch.randn(3, size, size, generator=g) * gain + off.view(3, 1, 1) + wave
    return out


DICT_INPUT = ("darknet53_det", "yolov3_darknet53", "mobilenet_v1_det")     # forward takes {"images": NCHW}


def model_input(name, x):
    return {"images": x} if name in DICT_INPUT else x


def flatten_outputs(y):
    """Outputs of a forward as a flat list of tensors: a tensor, a list of maps, or YOLOv3's dict (body, neck, head order)."""
    if isinstance(y, dict):
        return [t for k in ("body_feats", "neck_feats", "yolo_head_outs") for t in y[k]]
    return list(y) if isinstance(y, (list, tuple)) else [y]


def parity_stats(y, ref):
    """Logit parity of a CUDA result against the oracle, as the north star states it: max-abs error, top-1 agreement,
    and the context needed to read them (how many oracle decisions are closer than the error bound, how much the
    logits depend on the image at all)."""
    y, ref = y.detach().float().cpu(), ref.detach().float().cpu()
    err = (y - ref).abs()
    top2 = ref.topk(2, dim=1).values
    margin = top2[:, 0] - top2[:, 1]
    agree = y.argmax(1) == ref.argmax(1)
    max_abs = float(err.max())
    # a disagreement is only possible where the oracle's own margin is within twice the largest logit error
    explained = agree | (margin <= 2 * max_abs)
    return {
        "n": int(y.shape[0]), "max_abs": max_abs, "top1_agree": int(agree.sum()),
        "top1_disagree_explained_by_margin": int((~agree & explained).sum()),
        "top1_unexplained": int((~explained).sum()),
        "distinct_top1": len(set(ref.argmax(1).tolist())),
        "min_margin": float(margin.min()), "n_margin_below_2e-2": int((margin < 2e-2).sum()),
        "logit_std_over_classes": float(ref.std(1).mean()), "logit_std_over_images": float(ref.std(0).mean()) if y.shape[0] > 1 else 0.0,
    }


def state_dict_digest(sd) -> str:
    """Order-sensitive SHA-256 over names, shapes and raw fp32 bytes."""
    import hashlib

    h = hashlib.sha256()
    for k, v in sd.items():
        h.update(k.encode())
        h.update(str(tuple(v.shape)).encode())
        h.update(v.detach().cpu().contiguous().numpy().tobytes())
    return h.hexdigest()
