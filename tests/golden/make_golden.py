#!/usr/bin/env python
"""Mint the golden fixtures for the CNN-backbone forward path.

Runs the REFERENCE'S OWN model files (loaded by path from /root/reference,
unmodified, through oracle/ref_loader.py + oracle/tlx_compat.py) on the seeded
weights of tlxcv_b200/testing.py and seeded synthetic images, and writes

  tests/golden/<model>.npz   logits (or feature maps) + input spec + weight digest
  tests/golden/manifests.json  ordered state-dict manifest (name -> shape) per model

The reference has no tests or golden vectors of its own, so these
files are the pin.  They can only be (re)generated where /root/reference is
mounted; the GPU box uses the committed files.

    python tests/golden/make_golden.py
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import ref_loader  # noqa: E402
from tlxcv_b200.testing import (flatten_outputs, model_input, seeded_state_dict, state_dict_digest,  # noqa: E402
                                structured_images, synthetic_images)

# model -> (n_images, image size)
CASES = {
    "resnet50": (4, 224),
    "resnet18": (2, 224),
    "resnext50_32x4d": (2, 224),
    "mobilenet_v1": (2, 224),
    "mobilenet_v2": (2, 224),
    "darknet53_cls": (2, 224),
    "darknet53_det": (1, 64),
    "yolov3_darknet53": (1, 64),          # backbone + YOLOv3FPN + head output convs: 9 maps
    "mobilenet_v1_det": (2, 96),
    "resnet50_vd": (1, 128),              # segmentation backbone: AvgPool2d shortcut, dilation 2 / 4; four stage outputs
    "resnet18_vd": (2, 96),
    "resnest50": (2, 128),                # split attention (radix 2), avd / avg_down average pools, deep stem
}
MANIFEST_ONLY = ["resnet34", "resnet101", "wide_resnet50_2", "resnext50_64x4d", "resnest101"]

# BASELINE.json configurations at their stated size: fixture file -> (model, n_images, image size, subsample).
# Images are testing.structured_images (image i depends on (seed, i) only), so they are the FIRST n images of the full batch the GPU test feeds
# (ResNet-50: the whole bs256 batch; MobileNetV2 bs512 / ResNeXt-50 bs256: the first 64 images; DarkNet-53 bs64 at
# 608x608: the first 2 images).  DarkNet's three feature maps are 20 MB per image pair: the fixture keeps a strided
# subsample ([:, ::4, ::3, ::3]) plus float64 sum / abs-sum checksums of the full maps.
FULL_SIZE = {
    "resnet50_bs256": ("resnet50", 256, 224, False),
    "mobilenet_v2_bs64": ("mobilenet_v2", 64, 224, False),
    "resnext50_32x4d_bs64": ("resnext50_32x4d", 64, 224, False),
    "darknet53_det_608": ("darknet53_det", 2, 608, True),
    "yolov3_darknet53_608": ("yolov3_darknet53", 1, 608, True),
}
SUB = (slice(None), slice(None, None, 4), slice(None, None, 3), slice(None, None, 3))


def main():
    torch.set_num_threads(os.cpu_count() or 1)
    out_dir = os.path.dirname(os.path.abspath(__file__))
    manifests = {}
    only = sys.argv[1:]                      # `make_golden.py resnest50 resnest101`: (re)mint these only, keep the rest
    if only:
        with open(os.path.join(out_dir, "manifests.json")) as f:
            manifests = json.load(f)
    for name in list(CASES) + MANIFEST_ONLY:
        if only and name not in only:
            continue
        model = ref_loader.build(name)
        manifest = [(k, list(v.shape)) for k, v in model.state_dict().items()]
        manifests[name] = manifest
        if name not in CASES:
            continue
        n, size = CASES[name]
        sd = seeded_state_dict(model.state_dict(), name)
        model.load_state_dict(sd)
        model.set_eval()
        x = synthetic_images(n, size)
        with torch.no_grad():
            outs = flatten_outputs(model(model_input(name, x)))
        arrays = {f"out{i}": o.numpy() for i, o in enumerate(outs)}
        np.savez_compressed(
            os.path.join(out_dir, f"{name}.npz"),
            n=np.int64(n), size=np.int64(size),
            weight_digest=np.array(state_dict_digest(sd)),
            input_digest=np.array(state_dict_digest({"x": x})),
            **arrays,
        )
        print(name, [tuple(o.shape) for o in outs], "std", float(outs[0].std()))
    for fname, (name, n, size, sub) in FULL_SIZE.items():
        if only and fname not in only:
            continue
        model = ref_loader.build(name)
        sd = seeded_state_dict(model.state_dict(), name)
        model.load_state_dict(sd)
        model.set_eval()
        x = structured_images(n, size)
        with torch.no_grad():
            outs = flatten_outputs(model(model_input(name, x)))
        arrays = {}
        for i, o in enumerate(outs):
            arrays[f"out{i}"] = (o[SUB] if sub else o).contiguous().numpy()
            if sub:
                arrays[f"sum{i}"] = np.float64(o.double().sum().item())
                arrays[f"abssum{i}"] = np.float64(o.double().abs().sum().item())
                arrays[f"shape{i}"] = np.array(o.shape, dtype=np.int64)
        np.savez_compressed(
            os.path.join(out_dir, f"{fname}.npz"),
            model=np.array(name), n=np.int64(n), size=np.int64(size), subsampled=np.int64(1 if sub else 0),
            weight_digest=np.array(state_dict_digest(sd)),
            input_digest=np.array(state_dict_digest({"x": x})),
            **arrays,
        )
        print(fname, [tuple(o.shape) for o in outs], "std", float(outs[0].std()))
    with open(os.path.join(out_dir, "manifests.json"), "w") as f:
        json.dump(manifests, f)
    if only:
        print("wrote", out_dir, only)
        return
    # OpenCV INTER_LINEAR vectors for oracle/cv_resize.py (the arithmetic behind tensorlayerx's Resize on numpy images)
    import cv2
    rng = np.random.default_rng(0)
    vec = {"cv2_version": np.array(cv2.__version__)}
    for k, (hs, ws, h, w) in enumerate([(37, 53, 64, 96), (120, 90, 32, 32), (17, 9, 5, 3), (48, 64, 224, 224), (31, 31, 31, 31)]):
        img = rng.integers(0, 256, (hs, ws, 3), dtype=np.uint8)
        vec[f"src{k}"], vec[f"dst{k}"] = img, cv2.resize(img, (w, h), interpolation=cv2.INTER_LINEAR)
    np.savez_compressed(os.path.join(out_dir, "cv_resize.npz"), **vec)
    with open(os.path.join(out_dir, "manifests.json"), "w") as f:
        json.dump(manifests, f)
    print("wrote", out_dir)


if __name__ == "__main__":
    main()
