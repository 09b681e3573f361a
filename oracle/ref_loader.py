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
    "resnest": "tlxcv/models/classification/resnest.py",
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


_PKG = "_tlxcv_ref"


def load_reference_package_module(dotted: str, shim=None):
    """Import ``tlxcv.<dotted>`` (e.g. ``models.detection.yolov3``) from the reference tree WITH its relative imports
    (``from .backbones.darknet import ...``, ``from .utils.layers import Interpolater``) but without executing
    ``tlxcv/__init__.py`` / ``tlxcv/models/__init__.py`` (which star-import every family): the parent packages are
    empty namespace stand-ins whose ``__path__`` points into the reference tree."""
    import types

    parts = dotted.split(".")
    shim = shim or tlx_compat.installed
    created = []
    with shim():
        try:
            for i in range(len(parts)):
                name = ".".join([_PKG] + parts[:i])
                path = os.path.join(REFERENCE_ROOT, "tlxcv", *parts[:i])
                if not os.path.isdir(path):
                    raise FileNotFoundError(path)
                if name not in sys.modules:
                    m = types.ModuleType(name)
                    m.__path__, m.__package__ = [path], name
                    sys.modules[name] = m
                    created.append(name)
            return importlib.import_module(".".join([_PKG] + parts))
        finally:
            # the stand-in packages and everything imported under them go away again (the next shim must re-import)
            for k in [k for k in sys.modules if k == _PKG or k.startswith(_PKG + ".")]:
                del sys.modules[k]


def _build_yolov3(shim=None, **kw):
    """DarkNet-53 -> YOLOv3FPN -> YOLOv3Head of the reference (detection/yolov3.py:51-68 without the host-side
    post-processing): returns the raw head maps the way ``YOLOv3.forward`` computes them in eval mode."""
    y = load_reference_package_module("models.detection.yolov3", shim)
    shim = shim or tlx_compat.installed
    with shim():
        nn = sys.modules["tensorlayerx.nn"]

        class YOLOv3Net(nn.Module):
            def __init__(self):
                super().__init__()
                self.backbone = y.DarkNet()
                self.neck = y.YOLOv3FPN()
                self.yolo_head = y.YOLOv3Head(**kw)

            def forward(self, inputs):
                body_feats = self.backbone(inputs)
                neck_feats = self.neck(body_feats, False)
                return {"body_feats": body_feats, "neck_feats": neck_feats, "yolo_head_outs": self.yolo_head(neck_feats)}

        return YOLOv3Net()


def _build_resnet_vd(shim=None, **kw):
    """segmentation/backbones/resnet_vd.py with ITS OWN helper files (`from ..layers import Activation, Add`): the helpers
    are loaded one by one from segmentation/layers/activation.py and wrap_functions.py; the package's __init__ (which also
    pulls in the pyramid-pooling / JPU layers of the segmentation heads) is not executed."""
    import types

    shim = shim or tlx_compat.installed
    seg = os.path.join(REFERENCE_ROOT, "tlxcv", "models", "segmentation")
    with shim():
        created = []
        try:
            for name, path in ((_PKG, os.path.join(REFERENCE_ROOT, "tlxcv")), (_PKG + ".models", os.path.join(REFERENCE_ROOT, "tlxcv", "models")),
                               (_PKG + ".models.segmentation", seg), (_PKG + ".models.segmentation.backbones", os.path.join(seg, "backbones")),
                               (_PKG + ".models.segmentation.layers", os.path.join(seg, "layers"))):
                m = types.ModuleType(name)
                m.__path__, m.__package__ = [path], name
                sys.modules[name] = m
                created.append(name)
            layers = sys.modules[_PKG + ".models.segmentation.layers"]
            for fname, names in (("activation", ("Activation",)), ("wrap_functions", ("Add",))):
                sub = importlib.import_module(f"{_PKG}.models.segmentation.layers.{fname}")
                for n in names:
                    setattr(layers, n, getattr(sub, n))
            mod = importlib.import_module(_PKG + ".models.segmentation.backbones.resnet_vd")
            return mod.ResNet_vd(**kw)
        finally:
            for k in [k for k in sys.modules if k == _PKG or k.startswith(_PKG + ".")]:
                del sys.modules[k]


def _build_det_mobilenet(shim=None, **kw):
    m = load_reference_package_module("models.detection.backbones.mobilenet_v1", shim)
    shim = shim or tlx_compat.installed
    with shim():
        return m.MobileNet(**kw)


# constructor table: name -> (file key, callable name, kwargs)
MODELS = {
    "resnet18": ("resnet", "resnet18", {}),
    "resnet34": ("resnet", "resnet34", {}),
    "resnet50": ("resnet", "resnet50", {}),
    "resnet101": ("resnet", "resnet101", {}),
    "wide_resnet50_2": ("resnet", "wide_resnet50_2", {}),
    "resnext50_32x4d": ("resnext", "resnext50_32x4d", {}),
    "resnext50_64x4d": ("resnext", "resnext50_64x4d", {}),
    "resnest50": ("resnest", "resnest50", {}),
    "resnest50_fast_1s1x64d": ("resnest", "resnest50_fast_1s1x64d", {}),
    "resnest101": ("resnest", "resnest101", {}),
    "mobilenet_v1": ("mobilenetv1", "MobileNetV1", {}),
    "mobilenet_v2": ("mobilenetv2", "mobilenet_v2", {}),
    "darknet53_cls": ("darknet53", "darknet53", {}),
    "darknet53_det": ("det_darknet", "DarkNet", {}),
}


PACKAGE_MODELS = {"yolov3_darknet53": _build_yolov3, "mobilenet_v1_det": _build_det_mobilenet,
                  "resnet50_vd": _build_resnet_vd, "resnet18_vd": lambda shim=None, **kw: _build_resnet_vd(shim, layers=18, **kw)}


def build(name: str, shim=None, **kwargs):
    """Construct reference model ``name`` (see MODELS / PACKAGE_MODELS) from its own file."""
    if name in PACKAGE_MODELS:
        return PACKAGE_MODELS[name](shim, **kwargs)
    key, fn, kw = MODELS[name]
    mod = load_reference_module(key, shim)
    kw = dict(kw, **kwargs)
    shim = shim or tlx_compat.installed
    with shim():
        return getattr(mod, fn)(**kw)
