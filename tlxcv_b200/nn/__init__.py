"""The tensorlayerx layer API for the CNN-backbone forward path, B200-backed.

Mirror of the ``tensorlayerx.nn`` symbols the reference hot-path files resolve
(SURVEY.md §8(b)): same class names, constructor arguments, parameter leaf
names (``filters`` OIHW, ``biases``, ``beta``/``gamma``/``moving_mean``/
``moving_var``, ``weights`` (in, out)) and creation order, so state dicts and
positional ``.npz`` weight files interchange with the reference.

What differs is execution.  A layer never computes by itself: calling any
``Module`` with CUDA tensors traces its ``forward`` into a graph (graph.py),
fuses it (planner.py) and runs the fused plan through the C-ABI library on
sm_100a kernels (runtime.py).  There is no CPU path and no training path:
both raise.
"""
from __future__ import annotations

import numpy as np
import torch

from .. import graph as _g
from . import initializers
from .initializers import _resolve

__all__ = [
    "Module", "Layer", "Sequential", "GroupConv2d", "Conv2d", "BatchNorm", "BatchNorm2d", "BatchNorm2D", "ReLU",
    "ReLU6", "LeakyReLU", "Dropout", "MaxPool2d", "AvgPool2d", "AdaptiveAvgPool2d", "Linear", "initializers", "layers", "layer",
]

_ACT_NAMES = {"relu": ("relu", 0.0), "relu6": ("relu6", 0.0), "leaky_relu": ("leaky", 0.2), "lrelu": ("leaky", 0.2)}


class _LayerList(list):
    """Python list attribute whose Module items are registered on the owner.

    The reference keeps some sub-modules in plain lists
    (detection/backbones/darknet.py:270-271,285,297).  Items are registered as
    ``<attr>.<i>`` unless the same object is already a registered child
    (classification/resnext.py:176-190 appends setattr-registered blocks).
    """

    def __init__(self, owner, attr, items=()):
        super().__init__()
        self._owner, self._attr = owner, attr
        for it in items:
            self.append(it)

    def append(self, item):
        if type(item) is list and item and all(isinstance(v, torch.nn.Module) for v in item):
            # a list of lists (segmentation/backbones/resnet_vd.py:287 `self.stage_list.append(block_list)`): the inner list
            # registers as `<attr>.<i>.<j>`
            item = _NestedLayers(item)
        super().append(item)
        if isinstance(item, torch.nn.Module):
            mods = self._owner._modules
            if all(item is not m for m in mods.values()):
                if self._attr not in mods:
                    mods[self._attr] = torch.nn.ModuleList()
                mods[self._attr].append(item)

    def extend(self, items):
        for it in items:
            self.append(it)


class _NestedLayers(torch.nn.ModuleList):
    """Inner list of a list of lists of layers: iterates / indexes like the Python list it replaces."""


class Module(torch.nn.Module):
    """``tensorlayerx.nn.Module``: ``__init__(name=None, act=None)``, ``forward``,
    ``set_eval`` / ``set_train`` / ``is_train``, ``all_weights`` /
    ``trainable_weights``, ``save_weights`` / ``load_weights``."""

    def __init__(self, name=None, act=None):
        super().__init__()
        self.name = name
        self.is_train = True
        self._tlx_act = _parse_act(act)

    # -- structure ------------------------------------------------------------
    def __setattr__(self, key, value):
        if type(value) is list and "_modules" in self.__dict__ and all(isinstance(v, torch.nn.Module) for v in value):
            self._modules.pop(key, None)
            self.__dict__[key] = _LayerList(self, key, value)
            return
        super().__setattr__(key, value)

    # -- execution ------------------------------------------------------------
    def __call__(self, *args, **kwargs):
        g = _g.active()
        if g is not None:
            # a ResNeSt SplatConv (the reference's own class included) is lowered as ONE unit: its forward is tensor algebra
            # (split / add_n / reshape / transpose / softmax / multiply, classification/resnest.py:146-166) that the plan
            # runs as GAP + two 1x1 convs + one split-attention pass
            fwd = (lambda x: split_attention(self, x)) if _is_splat_conv(self) else self.forward
            name = g.names.get(id(self)) if hasattr(g, "names") else None
            if name:
                saved, g._path = g._path, [name]
                try:
                    return fwd(*args, **kwargs)
                finally:
                    g._path = saved
            return fwd(*args, **kwargs)
        from .. import runtime
        return runtime.run_module(self, args, kwargs)

    def forward(self, *args, **kwargs):
        raise NotImplementedError

    def _apply_act(self, x):
        if self._tlx_act is None:
            return x
        return _g.active().act(x, *self._tlx_act)

    # -- mode -----------------------------------------------------------------
    def _set_mode(self, train):
        for m in self.modules():
            if isinstance(m, Module):
                m.is_train = train
                m.__dict__.pop("_b200_plans", None)      # plans are traced per mode
        self.train(train)
        return self

    def set_eval(self):
        return self._set_mode(False)

    def set_train(self):
        return self._set_mode(True)

    # -- weights --------------------------------------------------------------
    @property
    def all_weights(self):
        return list(self.parameters())

    @property
    def trainable_weights(self):
        return [p for p in self.parameters() if p.requires_grad]

    @property
    def nontrainable_weights(self):
        return [p for p in self.parameters() if not p.requires_grad]

    def save_weights(self, file_path, format=None):
        """``.npz``: positional list over ``all_weights`` order; ``npz_dict``: by name
        (tensorlayerx ``Module.save_weights``, called at demo/image_classification/train-resnet.py:81)."""
        fmt = format or _format_of(file_path)
        if fmt == "npz":
            np.savez(file_path, params=np.array([p.detach().cpu().numpy() for p in self.all_weights], dtype=object))
        elif fmt == "npz_dict":
            np.savez(file_path, **{k: v.detach().cpu().numpy() for k, v in self.state_dict().items()})
        else:
            raise ValueError(f"unsupported weight format {fmt!r} (npz | npz_dict)")

    def invalidate_plans(self):
        """Drop every cached B200 plan under this module, so that the next call re-reads the parameters.

        Plans pack bf16 copies of the weights at build time and are rebuilt automatically when a parameter's
        ``(data_ptr, _version)`` changes - which in-place writes through ``p.data`` (``p.data.copy_(w)``, EMA updates) do
        NOT touch.  Call this after such updates."""
        for m in self.modules():
            m.__dict__.pop("_b200_plans", None)
        return self

    def load_weights(self, file_path, format=None, skip=False):
        fmt = format or _format_of(file_path)
        # the positional .npz format stores an object array of per-tensor arrays (needs pickle); named files do not
        data = np.load(file_path, allow_pickle=(fmt == "npz"))
        if fmt == "npz":
            arrays = list(data["params"])
            weights = self.all_weights
            if len(arrays) != len(weights):
                raise ValueError(f"{file_path}: {len(arrays)} arrays for {len(weights)} weights")
            with torch.no_grad():
                for p, a in zip(weights, arrays):
                    p.copy_(_fit(torch.from_numpy(np.asarray(a, dtype=np.float32)), p.shape))
            self.invalidate_plans()
        elif fmt == "npz_dict":
            sd = {k: torch.from_numpy(data[k]) for k in data.files}
            self.load_state_dict(sd, strict=not skip)
        else:
            raise ValueError(f"unsupported weight format {fmt!r} (npz | npz_dict)")
        return self


Layer = Module


def _format_of(path):
    p = str(path)
    if p.endswith(".npz"):
        return "npz"
    raise ValueError(f"cannot infer weight format from {p!r}; pass format='npz' or 'npz_dict'")


def _fit(t, shape):
    """Accept a Linear ``weights`` stored (out, in) instead of (in, out)."""
    if tuple(t.shape) == tuple(shape):
        return t
    if t.ndim == 2 and tuple(t.t().shape) == tuple(shape):
        return t.t()
    raise ValueError(f"weight shape {tuple(t.shape)} does not fit {tuple(shape)}")


def _parse_act(act):
    if act is None:
        return None
    if isinstance(act, str):
        key = act.lower()
        if key not in _ACT_NAMES:
            raise NotImplementedError(f"activation {act!r} is not on the B200 path")
        return _ACT_NAMES[key]
    if isinstance(act, type):
        act = act()
    if isinstance(act, ReLU6):
        return ("relu6", 0.0)
    if isinstance(act, ReLU):
        return ("relu", 0.0)
    if isinstance(act, LeakyReLU):
        return ("leaky", act.negative_slope)
    raise NotImplementedError(f"activation {act!r} is not on the B200 path")


def _pair(v, what):
    if isinstance(v, int):
        return (v, v)
    v = tuple(int(i) for i in v)
    if len(v) != 2:
        raise ValueError(f"{what} must be an int or a pair, got {v}")
    return v


def _check_format(data_format):
    if data_format not in ("channels_first", "NCHW"):
        raise NotImplementedError(
            "only data_format='channels_first' (the torch-backend convention, "
            "demo/image_classification/predict-resnet.py:37-39) is on the B200 path")


def _sym(x, who):
    if not isinstance(x, _g.SymTensor):
        raise RuntimeError(
            f"{who}.forward was reached with a {type(x).__name__}; tlxcv_b200 layers only execute inside a traced "
            "plan (call the module, do not call .forward directly)")
    return x


def _is_splat_conv(m):
    """Structural test for ResNeSt's split-attention conv (classification/resnest.py:84-143)."""
    return (type(m).__name__ == "SplatConv" and
            all(hasattr(m, a) for a in ("radix", "conv1", "conv2", "conv3", "rsoftmax", "avg_pool2d")))


def split_attention(m, x):
    """``SplatConv.forward`` (classification/resnest.py:146-166) on the plan.

    ``y = conv1(x)`` holds ``radix`` channel groups; the reference sums them, pools globally, runs ``conv2`` (+BN+ReLU) and
    ``conv3``, soft-maxes over the radix axis and returns the attention-weighted sum of the groups.  Here the global pool runs
    over all groups at once and the sum moves into ``conv2`` (filters repeated along C_in: mean and sum commute with the
    1x1 conv), then ONE pass does softmax x multiply x sum (``TLXCV_OP_SPLAT_APPLY``)."""
    g = _g.active()
    x = _sym(x, "SplatConv")
    radix = int(m.radix)
    if radix < 2:
        raise NotImplementedError("SplatConv with radix 1 gates with a sigmoid (resnest.py:78-79): not on the B200 path")
    y = m.conv1(x)
    pooled = g.gap(y)
    conv2 = [c for c in m.conv2.children() if isinstance(c, GroupConv2d)]
    bn2 = [c for c in m.conv2.children() if isinstance(c, BatchNorm)]
    if len(conv2) != 1 or len(bn2) != 1:
        raise NotImplementedError("SplatConv.conv2 must be one GroupConv2d followed by one BatchNorm")
    names = getattr(g, "names", {})
    saved, g._path = g._path, [names.get(id(conv2[0])) or ".".join(g._path)]
    try:
        h = conv2[0]._apply_act(g.conv(pooled, conv2[0], dup_in=radix))
    finally:
        g._path = saved
    logits = m.conv3(bn2[0](h))
    return g.splat_apply(y, logits, radix, int(m.rsoftmax.cardinality))


class Sequential(Module):
    """``nn.Sequential([l0, l1])`` and ``nn.Sequential(l0, l1)`` (resnet.py:247,284; ops_fusion.py:48)."""

    def __init__(self, *layers, name=None):
        super().__init__(name)
        if len(layers) == 1 and isinstance(layers[0], dict):     # OrderedDict of named layers (classification/resnest.py:479)
            for key, layer in layers[0].items():
                self.add_module(str(key), layer)
            return
        if len(layers) == 1 and isinstance(layers[0], (list, tuple)):
            layers = tuple(layers[0])
        for i, layer in enumerate(layers):
            self.add_module(str(i), layer)

    def __len__(self):
        return len(self._modules)

    def __iter__(self):
        return iter(self._modules.values())

    def __getitem__(self, idx):
        items = list(self._modules.values())
        return Sequential(items[idx]) if isinstance(idx, slice) else items[idx]

    def append(self, layer):
        self.add_module(str(len(self._modules)), layer)
        return self

    def forward(self, x):
        for layer in self._modules.values():
            x = layer(x)
        return x


class GroupConv2d(Module):
    """2-D (grouped) convolution; cross-correlation, zero padding, OIHW ``filters``.

    Every conv on the hot path is this class.  Bias exists only
    when ``b_init`` is truthy: the reference disables it with ``b_init=()``
    (resnet.py:43) or ``b_init=False`` (detection/backbones/darknet.py:45).
    """

    def __init__(self, out_channels=32, kernel_size=(1, 1), stride=(1, 1), act=None, padding="SAME",
                 data_format="channels_first", dilation=(1, 1), n_group=1, W_init="truncated_normal",
                 b_init="constant", in_channels=None, name=None):
        super().__init__(name, act)
        _check_format(data_format)
        if in_channels is None:
            raise NotImplementedError("GroupConv2d needs in_channels (lazy channel inference is not on the hot path)")
        kh, kw = _pair(kernel_size, "kernel_size")
        self.stride = _pair(stride, "stride")
        self.dilation = _pair(dilation, "dilation")
        if isinstance(padding, str):
            if padding.upper() == "VALID":
                padding = 0
            elif padding.upper() == "SAME" and self.stride == (1, 1):
                padding = (self.dilation[0] * (kh - 1) // 2, self.dilation[1] * (kw - 1) // 2)
            else:
                raise NotImplementedError("string padding other than VALID / stride-1 SAME; pass an int")
        self.padding = _pair(padding, "padding")
        self.n_group = int(n_group)
        self.in_channels, self.out_channels = int(in_channels), int(out_channels)
        if self.in_channels % self.n_group or self.out_channels % self.n_group:
            raise ValueError("in_channels and out_channels must be divisible by n_group")
        shape = (self.out_channels, self.in_channels // self.n_group, kh, kw)
        self.filters = torch.nn.Parameter(_resolve(W_init, "truncated_normal")(shape))
        if b_init:
            self.biases = torch.nn.Parameter(_resolve(b_init, "constant")((self.out_channels,)))
        else:
            self.biases = None

    def forward(self, x):
        return self._apply_act(_g.active().conv(_sym(x, "GroupConv2d"), self))


Conv2d = GroupConv2d


class BatchNorm(Module):
    """Batch normalisation, inference form: ``(x-mean)/sqrt(var+eps)*gamma+beta`` then ``act``.

    Parameter creation order beta, gamma, moving_mean, moving_var.  In a plan it
    is folded to an fp32 scale/shift applied in the conv epilogue (never into the
    bf16 weights).  Training mode (batch statistics) is not on the B200 path.
    """

    def __init__(self, decay=0.9, epsilon=1e-5, act=None, is_train=True, beta_init="zeros", gamma_init="ones",
                 moving_mean_init="zeros", moving_var_init="ones", num_features=None, data_format="channels_first",
                 name=None):
        super().__init__(name, act)
        _check_format(data_format)
        if num_features is None:
            raise NotImplementedError("BatchNorm needs num_features")
        self.decay, self.epsilon, self.num_features = decay, epsilon, int(num_features)
        shape = (self.num_features,)
        self.beta = torch.nn.Parameter(_resolve(beta_init, "zeros")(shape))
        self.gamma = torch.nn.Parameter(_resolve(gamma_init, "ones")(shape))
        self.moving_mean = torch.nn.Parameter(_resolve(moving_mean_init, "zeros")(shape), requires_grad=False)
        # tensorlayerx's own default is zeros [recalled]; ones keeps a fresh model finite
        # (the reference's xavier-uniform moving variance, resnext.py:49-50, can be negative: keep |.|)
        self.moving_var = torch.nn.Parameter(_resolve(moving_var_init, "ones")(shape).abs(), requires_grad=False)

    def forward(self, x):
        if self.is_train:
            raise NotImplementedError(
                "BatchNorm in training mode (batch statistics) is outside the B200 inference path; call set_eval()")
        return self._apply_act(_g.active().bn(_sym(x, "BatchNorm"), self))


BatchNorm2d = BatchNorm
BatchNorm2D = BatchNorm


class ReLU(Module):
    def forward(self, x):
        return _g.active().act(_sym(x, "ReLU"), "relu")


class ReLU6(Module):
    def forward(self, x):
        return _g.active().act(_sym(x, "ReLU6"), "relu6")


class LeakyReLU(Module):
    def __init__(self, negative_slope=0.01, name=None):
        super().__init__(name)
        self.negative_slope = float(negative_slope)

    def forward(self, x):
        return _g.active().act(_sym(x, "LeakyReLU"), "leaky", self.negative_slope)


class Dropout(Module):
    """Identity in eval mode (classification/mobilenetv2.py:98)."""

    def __init__(self, p=0.5, seed=0, name=None):
        super().__init__(name)
        self.p = p

    def forward(self, x):
        if self.is_train:
            raise NotImplementedError("Dropout in training mode is outside the B200 inference path; call set_eval()")
        return x


class MaxPool2d(Module):
    def __init__(self, kernel_size=(3, 3), stride=(2, 2), padding="SAME", return_mask=False,
                 data_format="channels_first", name=None):
        super().__init__(name)
        _check_format(data_format)
        if return_mask:
            raise NotImplementedError("MaxPool2d(return_mask=True)")
        self.kernel_size = _pair(kernel_size, "kernel_size")
        self.stride = _pair(stride, "stride")
        if isinstance(padding, str):
            if padding.upper() != "VALID":
                raise NotImplementedError("MaxPool2d string padding other than VALID; pass an int")
            padding = 0
        self.padding = _pair(padding, "padding")

    def forward(self, x):
        return _g.active().maxpool(_sym(x, "MaxPool2d"), self.kernel_size, self.stride, self.padding)


class AvgPool2d(Module):
    """``nn.AvgPool2d(kernel_size, stride, padding)``; ``padding="SAME"`` is accepted where it adds nothing (windows that tile
    the map exactly: the 2x2 / stride-2 pool of ResNet_vd's shortcut on even maps, segmentation/backbones/resnet_vd.py:25-27)."""

    def __init__(self, kernel_size=(2, 2), stride=(2, 2), padding="SAME", data_format="channels_first", ceil_mode=False,
                 name=None):
        super().__init__(name)
        _check_format(data_format)
        # ceil_mode (classification/resnest.py:270-276, the dilated variants' 1x1 / stride-1 pool) only matters where the
        # windows do not tile the map, which the trace refuses anyway
        self.ceil_mode = bool(ceil_mode)
        self.kernel_size = _pair(kernel_size, "kernel_size")
        self.stride = _pair(stride, "stride")
        if isinstance(padding, str):
            if padding.upper() not in ("SAME", "VALID"):
                raise NotImplementedError(f"AvgPool2d padding {padding!r}")
            padding = 0          # SAME on a map the windows tile exactly == no padding (checked when traced)
        self.padding = _pair(padding, "padding")

    def forward(self, x):
        return _g.active().avgpool(_sym(x, "AvgPool2d"), self.kernel_size, self.stride, self.padding)


class AdaptiveAvgPool2d(Module):
    def __init__(self, output_size, data_format="channels_first", name=None):
        super().__init__(name)
        _check_format(data_format)
        if _pair(output_size, "output_size") != (1, 1):
            raise NotImplementedError("AdaptiveAvgPool2d: only global pooling (output_size=1) is on the hot path")

    def forward(self, x):
        return _g.active().gap(_sym(x, "AdaptiveAvgPool2d"))


class Linear(Module):
    """``y = x @ weights + biases`` with ``weights`` stored (in_features, out_features)."""

    def __init__(self, out_features, act=None, W_init="truncated_normal", b_init="constant", in_features=None,
                 name=None):
        super().__init__(name, act)
        if in_features is None:
            raise NotImplementedError("Linear needs in_features")
        self.in_features, self.out_features = int(in_features), int(out_features)
        self.weights = torch.nn.Parameter(_resolve(W_init, "truncated_normal")((self.in_features, self.out_features)))
        if b_init is None or b_init is False or (isinstance(b_init, tuple) and not b_init):
            self.biases = None
        else:
            self.biases = torch.nn.Parameter(_resolve(b_init, "constant")((self.out_features,)))

    def _load_from_state_dict(self, state_dict, prefix, *args, **kwargs):
        key = prefix + "weights"
        if key in state_dict and self.in_features != self.out_features:
            state_dict[key] = _fit(state_dict[key], self.weights.shape)
        return super()._load_from_state_dict(state_dict, prefix, *args, **kwargs)

    def forward(self, x):
        if self._tlx_act is not None:
            raise NotImplementedError("Linear(act=...) is not on the hot path")
        return _g.active().linear(_sym(x, "Linear"), self)


class _ActivationNamespace:
    """``tensorlayerx.nn.layers.activation`` / ``nn.layer.activation`` as far as the reference reaches into it
    (segmentation/layers/activation.py:24-35 looks activations up by lower-cased class name and instantiates them)."""
    ReLU, ReLU6, LeakyReLU = ReLU, ReLU6, LeakyReLU


class _LayersNamespace:
    activation = _ActivationNamespace


layers = _LayersNamespace
layer = _LayersNamespace
