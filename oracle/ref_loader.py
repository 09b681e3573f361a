"""Execute the reference's OWN model files, unmodified, against ``tlx_compat``.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Files are loaded by path with
``importlib`` so that ``tlxcv/models/__init__.py:1-7`` (which star-imports
every family, several of which need ``paddle``) is never executed
(SURVEY.md §3.1).  Only usable where the reference tree is mounted
(``TLXCV_REFERENCE`` env var, default ``/root/reference``); never on the GPU box.
"""
from __future__ import annotations

import importlib.util
import os
import sys

from . import tlx_compat

REFERENCE_ROOT = os.environ.get("TLXCV_REFERENCE", "/root/reference")

_FILES = {
    "resnet": "tlxcv/models/classification/resnet.py",
    "resnext": "tlxcv/models/classification/resnext.py",
    "mobilenetv1": "tlxcv/models/classification/mobilenetv1.py",
    "mobilenetv2": "tlxcv/models/classification/mobilenetv2.py",
    "darknet53": "tlxcv/models/classification/darknet53.py",
    "det_darknet": "tlxcv/models/detection/backbones/darknet.py",
    "image_classification": "tlxcv/tasks/image_classification.py",
}


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, _FILES["resnet"]))


def load_reference_module(key: str, shim=None):
    """Import one reference file by path.

    ``shim`` is a context manager that exposes *some* implementation of the
    tensorlayerx API as ``sys.modules['tensorlayerx']`` while the file executes;
    the default is the CPU oracle stand-in.  (tests/test_dropin.py passes the
    product's own shim to prove the reference files run on it unmodified.)
    """
    path = os.path.join(REFERENCE_ROOT, _FILES[key])
    if not os.path.isfile(path):
        raise FileNotFoundError(path)
    cls_dir = os.path.join(REFERENCE_ROOT, "tlxcv/models/classification")
    shim = shim or tlx_compat.installed
    with shim():
        # classification/mobilenetv2.py:6-7 uses bare `from utils...` / `from ops...`
        sys.path.insert(0, cls_dir)
        saved = {k: sys.modules.pop(k) for k in list(sys.modules)
                 if k in ("utils", "ops") or k.startswith(("utils.", "ops."))}
        try:
            spec = importlib.util.spec_from_file_location(f"_tlxcv_ref_{key}", path)
            mod = importlib.util.module_from_spec(spec)
            spec.loader.exec_module(mod)
        finally:
            sys.path.remove(cls_dir)
            for k in [k for k in sys.modules if k in ("utils", "ops") or k.startswith(("utils.", "ops."))]:
                del sys.modules[k]
            sys.modules.update(saved)
    return mod


# constructor table: name -> (file key, callable name, kwargs)
MODELS = {
    "resnet18": ("resnet", "resnet18", {}),
    "resnet34": ("resnet", "resnet34", {}),
    "resnet50": ("resnet", "resnet50", {}),
    "resnet101": ("resnet", "resnet101", {}),
    "wide_resnet50_2": ("resnet", "wide_resnet50_2", {}),
    "resnext50_32x4d": ("resnext", "resnext50_32x4d", {}),
    "resnext50_64x4d": ("resnext", "resnext50_64x4d", {}),
    "mobilenet_v1": ("mobilenetv1", "MobileNetV1", {}),
    "mobilenet_v2": ("mobilenetv2", "mobilenet_v2", {}),
    "darknet53_cls": ("darknet53", "darknet53", {}),
    "darknet53_det": ("det_darknet", "DarkNet", {}),
}


def build(name: str, shim=None, **kwargs):
    """Construct reference model ``name`` (see MODELS) from its own file."""
    key, fn, kw = MODELS[name]
    mod = load_reference_module(key, shim)
    kw = dict(kw, **kwargs)
    shim = shim or tlx_compat.installed
    with shim():
        return getattr(mod, fn)(**kw)
