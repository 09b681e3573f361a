"""torch-CPU-fp32 stand-in for the tensorlayerx symbols on the hot path.

TEST INFRASTRUCTURE (see oracle/__init__.py).  Restates, per op, what
``tensorlayerx`` with ``TL_BACKEND=torch`` lowers to:

* ``GroupConv2d``  -> ``F.conv2d`` (cross-correlation, zero pad, OIHW ``filters``,
  ``groups=n_group``); bias only when ``b_init`` is truthy
  (reference call sites: classification/resnet.py:37-45,99-133,199-207;
  resnext.py:31-41; ops/ops_fusion.py:39-42; detection/backbones/darknet.py:38-47)
* ``BatchNorm`` / ``BatchNorm2d`` -> ``F.batch_norm(training=is_train, eps=1e-5)``
  then the optional ``act`` (resnext.py:46-52, darknet53.py:30-33)
* ``MaxPool2d`` -> ``F.max_pool2d`` (pads with -inf) (resnet.py:213-218)
* ``AdaptiveAvgPool2d`` -> ``F.adaptive_avg_pool2d`` (resnet.py:228-231)
* ``Linear`` -> ``x @ weights + biases`` with ``weights`` stored (in, out)
* ``ReLU`` / ``ReLU6`` / ``LeakyReLU(0.1)`` / ``Dropout`` (identity in eval)
* functional ``add, relu, reshape, flatten, argmax, get_tensor_shape,
  FlattenReshape, ops.squeeze``

Sub-modules that the reference keeps in plain Python lists
(detection/backbones/darknet.py:270-271,285,297) are registered under
``<attr>.<i>`` unless the same module object is already registered under
another name (classification/resnext.py:176-190 appends setattr-registered
blocks to ``block_list``).  See SURVEY.md §8(b) "Known ambiguity".
"""
from __future__ import annotations

import contextlib
import sys
import types

import torch
import torch.nn.functional as F

BN_EPS = 1e-5


# --------------------------------------------------------------------------- #
# Module base
# --------------------------------------------------------------------------- #
class _TrackedList(list):
    """A python list that registers appended modules on its owner."""

    def __init__(self, owner, attr, items=()):
        super().__init__()
        self._owner = owner
        self._attr = attr
        for it in items:
            self.append(it)

    def append(self, item):
        if type(item) is list and item and all(isinstance(v, torch.nn.Module) for v in item):
            item = torch.nn.ModuleList(item)      # list of lists (segmentation/backbones/resnet_vd.py:287): <attr>.<i>.<j>
        super().append(item)
        if isinstance(item, torch.nn.Module):
            owner = self._owner
            if not any(item is m for m in owner._modules.values()):
                holder = owner._modules.get(self._attr)
                if holder is None:
                    holder = torch.nn.ModuleList()
                    owner._modules[self._attr] = holder
                holder.append(item)


class Module(torch.nn.Module):
    def __init__(self, name=None, act=None):
        super().__init__()
        self.name = name
        self.is_train = True
        self._act_name = act if isinstance(act, str) else None

    def __setattr__(self, key, value):
        if type(value) is list and "_modules" in self.__dict__ and not key.startswith("_act"):
            if all(isinstance(v, torch.nn.Module) for v in value):
                self.__dict__[key] = _TrackedList(self, key, value)
                return
        super().__setattr__(key, value)

    def set_eval(self):
        for m in self.modules():
            if isinstance(m, Module):
                m.is_train = False
        self.eval()

    def set_train(self):
        for m in self.modules():
            if isinstance(m, Module):
                m.is_train = True
        self.train()

    def _post_act(self, x):
        if self._act_name is None:
            return x
        if self._act_name == "relu":
            return F.relu(x)
        raise NotImplementedError(self._act_name)

    @property
    def all_weights(self):
        return list(self.parameters())

    @property
    def trainable_weights(self):
        return [p for p in self.parameters() if p.requires_grad]


class Sequential(Module):
    def __init__(self, *layers):
        super().__init__()
        if len(layers) == 1 and isinstance(layers[0], dict):     # OrderedDict of named layers (classification/resnest.py:479)
            for k, l in layers[0].items():
                self.add_module(str(k), l)
            return
        if len(layers) == 1 and isinstance(layers[0], (list, tuple)):
            layers = tuple(layers[0])
        for i, l in enumerate(layers):
            self.add_module(str(i), l)

    def __len__(self):
        return len(self._modules)

    def __getitem__(self, i):
        return list(self._modules.values())[i]

    def forward(self, x):
        for m in self._modules.values():
            x = m(x)
        return x


def _pair(v):
    return (v, v) if isinstance(v, int) else tuple(v)


def _param(shape, requires_grad=True, fill=None):
    t = torch.zeros(shape) if fill is None else torch.full(shape, float(fill))
    return torch.nn.Parameter(t, requires_grad=requires_grad)


class GroupConv2d(Module):
    def __init__(self, out_channels=32, kernel_size=(1, 1), stride=(1, 1), act=None, padding="SAME",
                 data_format="channels_last", dilation=(1, 1), n_group=1, W_init="truncated_normal",
                 b_init="constant", in_channels=None, name=None):
        super().__init__(name, act)
        assert data_format == "channels_first", "oracle exercises the torch convention only"
        kh, kw = _pair(kernel_size)
        self.stride, self.dilation = _pair(stride), _pair(dilation)
        self.padding = _pair(padding)
        self.n_group = n_group
        self.filters = _param((out_channels, in_channels // n_group, kh, kw))
        with torch.no_grad():
            torch.nn.init.kaiming_normal_(self.filters)
        self.biases = _param((out_channels,)) if b_init else None

    def forward(self, x):
        y = F.conv2d(x, self.filters, self.biases, self.stride, self.padding, self.dilation, self.n_group)
        return self._post_act(y)


Conv2d = GroupConv2d


class BatchNorm(Module):
    def __init__(self, decay=0.9, epsilon=BN_EPS, act=None, is_train=True, beta_init="zeros", gamma_init="ones",
                 moving_mean_init="zeros", moving_var_init="zeros", num_features=None, data_format="channels_last",
                 name=None):
        super().__init__(name, act)
        self.epsilon = epsilon
        self.decay = decay
        self.beta = _param((num_features,))
        self.gamma = _param((num_features,), fill=1.0)
        self.moving_mean = _param((num_features,), requires_grad=False)
        self.moving_var = _param((num_features,), requires_grad=False, fill=1.0)

    def forward(self, x):
        y = F.batch_norm(x, self.moving_mean.data, self.moving_var.data, self.gamma, self.beta,
                         training=self.is_train, momentum=1.0 - self.decay, eps=self.epsilon)
        return self._post_act(y)


BatchNorm2d = BatchNorm
BatchNorm2D = BatchNorm


class ReLU(Module):
    def forward(self, x):
        return F.relu(x)


class ReLU6(Module):
    def forward(self, x):
        return F.relu6(x)


class LeakyReLU(Module):
    def __init__(self, negative_slope=0.01, name=None):
        super().__init__(name)
        self.negative_slope = negative_slope

    def forward(self, x):
        return F.leaky_relu(x, self.negative_slope)


class Dropout(Module):
    def __init__(self, p=0.5, seed=0, name=None):
        super().__init__(name)
        self.p = p

    def forward(self, x):
        return F.dropout(x, self.p, training=self.is_train)


class MaxPool2d(Module):
    def __init__(self, kernel_size=(3, 3), stride=(2, 2), padding="SAME", return_mask=False,
                 data_format="channels_last", name=None):
        super().__init__(name)
        self.kernel_size, self.stride, self.padding = kernel_size, stride, padding

    def forward(self, x):
        return F.max_pool2d(x, self.kernel_size, self.stride, self.padding)


class AvgPool2d(Module):
    """``padding="SAME"`` pads only where the windows do not tile the map (ceil mode, zeros excluded from the count is what
    tensorlayerx's torch backend does [recalled]); the hot path only meets maps the 2x2 / stride-2 windows tile exactly."""

    def __init__(self, kernel_size=(2, 2), stride=(2, 2), padding="SAME", data_format="channels_last", ceil_mode=False,
                 name=None):
        super().__init__(name)
        self.kernel_size, self.stride, self.padding, self.ceil_mode = kernel_size, stride, padding, ceil_mode

    def forward(self, x):
        k = self.kernel_size if isinstance(self.kernel_size, int) else self.kernel_size[0]
        st = self.stride if isinstance(self.stride, int) else self.stride[0]
        if isinstance(self.padding, str):
            if (x.shape[2] - k) % st or (x.shape[3] - k) % st:
                raise NotImplementedError("AvgPool2d(padding='SAME') on a map the windows do not tile")
            return F.avg_pool2d(x, k, st, 0)
        return F.avg_pool2d(x, k, st, self.padding, ceil_mode=bool(self.ceil_mode))


class AdaptiveAvgPool2d(Module):
    def __init__(self, output_size, data_format="channels_last", name=None):
        super().__init__(name)
        self.output_size = output_size

    def forward(self, x):
        return F.adaptive_avg_pool2d(x, self.output_size)


class Linear(Module):
    def __init__(self, out_features, act=None, W_init="truncated_normal", b_init="constant", in_features=None,
                 name=None):
        super().__init__(name, act)
        self.weights = _param((in_features, out_features))
        with torch.no_grad():
            torch.nn.init.normal_(self.weights, std=in_features ** -0.5)
        self.biases = _param((out_features,)) if (b_init is not None and b_init is not False) else None

    def forward(self, x):
        y = torch.matmul(x, self.weights)
        if self.biases is not None:
            y = y + self.biases
        return self._post_act(y)


class FlattenReshape(Module):
    def forward(self, x):
        return x.reshape(x.shape[0], -1)


# functional ----------------------------------------------------------------- #
def add(value, bias):
    return torch.add(value, bias)


def relu(x):
    return F.relu(x)


def reshape(tensor, shape):
    return torch.reshape(tensor, tuple(shape))


def flatten(x, start_axis=0, stop_axis=-1):
    return torch.flatten(x, start_axis, stop_axis)


def squeeze(x, axis=None):
    if axis is None:
        return torch.squeeze(x)
    for a in sorted(([axis] if isinstance(axis, int) else list(axis)), reverse=True):
        x = torch.squeeze(x, a)
    return x


def argmax(x, axis=None):
    return torch.argmax(x, dim=axis)


def get_tensor_shape(x):
    return list(x.shape)


def concat(values, axis=0):
    return torch.cat(list(values), dim=axis)


def split(value, num_or_size_splits, axis=0):
    """classification/resnest.py:150-151,160-161: equal parts (int) or the given sizes along ``axis``."""
    if isinstance(num_or_size_splits, int):
        return list(torch.chunk(value, num_or_size_splits, dim=axis))
    return list(torch.split(value, list(num_or_size_splits), dim=axis))


def add_n(inputs):
    out = inputs[0]
    for t in inputs[1:]:
        out = out + t
    return out


def multiply(x, y):
    return torch.mul(x, y)


def transpose(a, perm=None, conjugate=False):
    return a.permute(*perm) if perm is not None else a.t()


def softmax(logits, axis=-1):
    return F.softmax(logits, dim=axis)


def sigmoid(x):
    return torch.sigmoid(x)


class _Init:
    """Inert initializer: weights are injected by the harness."""

    def __init__(self, *a, **k):
        pass

    def __bool__(self):
        return True


# --------------------------------------------------------------------------- #
# sys.modules installation
# --------------------------------------------------------------------------- #
def _build_modules():
    tlx = types.ModuleType("tensorlayerx")
    nn = types.ModuleType("tensorlayerx.nn")
    init = types.ModuleType("tensorlayerx.nn.initializers")
    ops = types.ModuleType("tensorlayerx.ops")
    for name in ("xavier_uniform", "random_uniform", "truncated_normal", "constant", "ones", "zeros",
                 "random_normal", "he_normal"):
        setattr(init, name, _Init)
    for cls in (Module, Sequential, GroupConv2d, Conv2d, BatchNorm, BatchNorm2d, ReLU, ReLU6, LeakyReLU, Dropout,
                MaxPool2d, AvgPool2d, AdaptiveAvgPool2d, Linear):
        setattr(nn, cls.__name__, cls)
    # nn.layers.activation / nn.layer.activation: segmentation/layers/activation.py:24-35 looks activations up by name
    act_ns = types.ModuleType("tensorlayerx.nn.layers.activation")
    act_ns.ReLU, act_ns.ReLU6, act_ns.LeakyReLU = ReLU, ReLU6, LeakyReLU
    layers_ns = types.ModuleType("tensorlayerx.nn.layers")
    layers_ns.activation = act_ns
    nn.layers = nn.layer = layers_ns
    nn.BatchNorm2d = BatchNorm
    nn.BatchNorm2D = BatchNorm
    nn.Conv2d = GroupConv2d
    nn.Layer = Module
    nn.initializers = init
    for fn in (add, relu, reshape, flatten, squeeze, argmax, get_tensor_shape, concat, split, add_n, multiply, transpose, softmax,
               sigmoid):
        setattr(tlx, fn.__name__, fn)
        setattr(ops, fn.__name__, fn)
    tlx.FlattenReshape = FlattenReshape
    tlx.ReLU = ReLU
    tlx.BACKEND = "torch"
    tlx.nn = nn
    tlx.ops = ops
    tlx.initializers = init
    mods = {
        "tensorlayerx": tlx, "tensorlayerx.nn": nn, "tensorlayerx.nn.initializers": init,
        "tensorlayerx.ops": ops, "tensorlayerx.initializers": init,
    }
    # inert paddle / paddle2tlx stubs for the two paddle2tlx-converted files
    # (classification/mobilenetv2.py:2-8, classification/darknet53.py:2-13)
    paddle = types.ModuleType("paddle")
    p2t = types.ModuleType("paddle2tlx")
    pd2 = types.ModuleType("paddle2tlx.pd2tlx")
    utils = types.ModuleType("paddle2tlx.pd2tlx.utils")
    pops = types.ModuleType("paddle2tlx.pd2tlx.ops")
    tlxops = types.ModuleType("paddle2tlx.pd2tlx.ops.tlxops")

    def restore_model_clas(model, *a, **k):
        raise RuntimeError("pretrained weights are unavailable offline")

    utils.restore_model_clas = restore_model_clas
    tlxops.tlx_Dropout = Dropout
    p2t.pd2tlx = pd2
    pd2.utils, pd2.ops = utils, pops
    pops.tlxops = tlxops
    mods.update({
        "paddle": paddle, "paddle2tlx": p2t, "paddle2tlx.pd2tlx": pd2, "paddle2tlx.pd2tlx.utils": utils,
        "paddle2tlx.pd2tlx.ops": pops, "paddle2tlx.pd2tlx.ops.tlxops": tlxops,
    })
    return mods


@contextlib.contextmanager
def installed():
    """Temporarily expose the stand-in as ``tensorlayerx`` (+ paddle stubs)."""
    mods = _build_modules()
    saved = {k: sys.modules.get(k) for k in mods}
    sys.modules.update(mods)
    try:
        yield mods["tensorlayerx"]
    finally:
        for k, v in saved.items():
            if v is None:
                sys.modules.pop(k, None)
            else:
                sys.modules[k] = v
