"""tlxcv_b200 — B200-native CNN-backbone forward path for TLXCV models.

The package plays two roles for the one hot path it covers
(BASELINE.json ``north_star``):

* ``tlxcv_b200`` / ``tlxcv_b200.nn`` expose the ``tensorlayerx`` symbols the
  reference's model files use (``import tlxcv_b200 as tlx``), and
  ``install_as_tensorlayerx()`` registers the package under that name so the
  reference's own files run on it unmodified with ``TL_BACKEND=torch``;
* ``tlxcv_b200.models`` / ``tlxcv_b200.tasks`` mirror the ``tlxcv.models``
  constructors and ``tlxcv.tasks.ImageClassification`` on the hot path.

Execution is always: trace -> fuse -> one C-ABI call -> sm_100a kernels
(``libtlxcv_b200.so``).  There is no CPU or eager fallback.
"""
from __future__ import annotations

import contextlib
import sys

from . import graph as _g
from . import nn
from .nn import Module  # noqa: F401

BACKEND = "torch"
__version__ = "0.1.0"


# --------------------------------------------------------------------------- #
# functional API (tensorlayerx top level)
# --------------------------------------------------------------------------- #
def _traced(x, who):
    if not isinstance(x, _g.SymTensor):
        raise RuntimeError(f"tlxcv_b200.{who} is only defined inside a traced Module forward (no eager path)")
    return x.graph


def add(value, bias):
    """``tlx.add(value, bias)`` (resnext.py:117, detection/backbones/darknet.py:158)."""
    return _traced(value, "add").add(value, bias)


def relu(x):
    return _traced(x, "relu").act(x, "relu")


def relu6(x):
    return _traced(x, "relu6").act(x, "relu6")


def leaky_relu(x, negative_slope=0.01):
    return _traced(x, "leaky_relu").act(x, "leaky", negative_slope)


def reshape(tensor, shape):
    return _traced(tensor, "reshape").reshape(tensor, shape)


def flatten(x, start_axis=0, stop_axis=-1):
    """``tensorlayerx.flatten(x, 1)`` (mobilenetv2.py:107)."""
    if start_axis != 1 or stop_axis not in (-1, len(x.shape) - 1):
        raise NotImplementedError("flatten: only flatten(x, 1) is on the hot path")
    return _traced(x, "flatten").reshape(x, (x.shape[0], -1))


def squeeze(x, axis=None):
    """``tensorlayerx.ops.squeeze(x, axis=[2, 3])`` (darknet53.py:131)."""
    g = _traced(x, "squeeze")
    axes = range(len(x.shape)) if axis is None else ([axis] if isinstance(axis, int) else list(axis))
    axes = {a % len(x.shape) for a in axes}
    for a in axes:
        if axis is not None and x.shape[a] != 1:
            raise ValueError(f"squeeze: axis {a} has size {x.shape[a]}")
    shape = [s for i, s in enumerate(x.shape) if not (i in axes and s == 1)]
    return g.reshape(x, shape)


def argmax(x, axis=None):
    """``tlx.argmax(outputs, axis=-1)`` (tasks/image_classification.py:23)."""
    return _traced(x, "argmax").argmax(x, -1 if axis is None else axis)


def concat(values, axis=0):
    """``tlx.concat([route, x], axis=1)`` (detection/yolov3.py:244)."""
    values = list(values)
    return _traced(values[0], "concat").concat(values, axis)


def softmax(logits, axis=-1):
    """``tlx.softmax`` over the class axis of the fp32 logits (the "final FC + softmax" of the north star)."""
    return _traced(logits, "softmax").softmax(logits, axis)


class _Losses:
    """``tensorlayerx.losses`` (only what ``ImageClassification.loss_fn`` uses, tasks/image_classification.py:10-15)."""

    @staticmethod
    def softmax_cross_entropy_with_logits(output, target, reduction="mean"):
        if reduction != "mean":
            raise NotImplementedError("softmax_cross_entropy_with_logits: only reduction='mean'")
        return _traced(output, "losses.softmax_cross_entropy_with_logits").softmax_ce(output, target)


losses = _Losses()


def get_tensor_shape(x):
    return list(x.shape)


class FlattenReshape(nn.Module):
    """``tlx.FlattenReshape()`` (resnet.py:232)."""

    def forward(self, x):
        return _traced(x, "FlattenReshape").reshape(x, (x.shape[0], -1))


ReLU = nn.ReLU


class _Ops:
    """``tensorlayerx.ops`` namespace (only ``squeeze`` is used, darknet53.py:131)."""
    squeeze = staticmethod(squeeze)
    add = staticmethod(add)
    relu = staticmethod(relu)
    reshape = staticmethod(reshape)
    flatten = staticmethod(flatten)
    argmax = staticmethod(argmax)
    concat = staticmethod(concat)


ops = _Ops()
initializers = nn.initializers


# --------------------------------------------------------------------------- #
# drop-in shim
# --------------------------------------------------------------------------- #
_SHIM_NAMES = ("tensorlayerx", "tensorlayerx.nn", "tensorlayerx.nn.initializers", "tensorlayerx.ops",
               "tensorlayerx.initializers")


def _shim_modules():
    import types

    me = sys.modules[__name__]
    ops_mod = types.ModuleType("tensorlayerx.ops")
    for k in ("squeeze", "add", "relu", "reshape", "flatten", "argmax", "concat"):
        setattr(ops_mod, k, getattr(_Ops, k))
    return {"tensorlayerx": me, "tensorlayerx.nn": nn, "tensorlayerx.nn.initializers": nn.initializers,
            "tensorlayerx.ops": ops_mod, "tensorlayerx.initializers": nn.initializers}


def install_as_tensorlayerx():
    """Register this package as ``tensorlayerx`` so that reference model files
    (``import tensorlayerx as tlx; import tensorlayerx.nn as nn``) build B200-backed
    modules without modification.  See INTEGRATION.md."""
    if "tensorlayerx" in sys.modules and sys.modules["tensorlayerx"] is not sys.modules[__name__]:
        raise RuntimeError("a different tensorlayerx is already imported")
    sys.modules.update(_shim_modules())


@contextlib.contextmanager
def as_tensorlayerx():
    """Scoped variant of :func:`install_as_tensorlayerx` (used by the drop-in tests)."""
    saved = {k: sys.modules.get(k) for k in _SHIM_NAMES}
    sys.modules.update(_shim_modules())
    try:
        yield sys.modules[__name__]
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v


from . import models, tasks  # noqa: E402,F401
