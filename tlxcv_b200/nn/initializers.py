"""``tensorlayerx.nn.initializers`` subset used by the hot-path files
(resnext.py:6,39,49-50,196; darknet53.py:6,12,27-32,97)."""
from __future__ import annotations

import math

import torch

__all__ = ["Initializer", "zeros", "ones", "constant", "random_uniform", "random_normal", "truncated_normal",
           "xavier_uniform", "he_normal"]


class Initializer:
    def __call__(self, shape, dtype=torch.float32):
        raise NotImplementedError


def _fans(shape):
    if len(shape) == 1:
        return shape[0], shape[0]
    if len(shape) == 2:
        return shape[0], shape[1]
    rf = 1
    for s in shape[2:]:
        rf *= s
    return shape[1] * rf, shape[0] * rf      # OIHW


class constant(Initializer):
    def __init__(self, value=0.0):
        self.value = value

    def __call__(self, shape, dtype=torch.float32):
        return torch.full(tuple(shape), float(self.value), dtype=dtype)


class zeros(constant):
    def __init__(self):
        super().__init__(0.0)


class ones(constant):
    def __init__(self):
        super().__init__(1.0)


class random_uniform(Initializer):
    def __init__(self, minval=-0.05, maxval=0.05, seed=None):
        self.minval, self.maxval = minval, maxval

    def __call__(self, shape, dtype=torch.float32):
        return torch.empty(tuple(shape), dtype=dtype).uniform_(self.minval, self.maxval)


class random_normal(Initializer):
    def __init__(self, mean=0.0, stddev=0.05, seed=None):
        self.mean, self.stddev = mean, stddev

    def __call__(self, shape, dtype=torch.float32):
        return torch.empty(tuple(shape), dtype=dtype).normal_(self.mean, self.stddev)


class truncated_normal(Initializer):
    def __init__(self, mean=0.0, stddev=0.02, seed=None):
        self.mean, self.stddev = mean, stddev

    def __call__(self, shape, dtype=torch.float32):
        t = torch.empty(tuple(shape), dtype=dtype)
        return torch.nn.init.trunc_normal_(t, self.mean, self.stddev, self.mean - 2 * self.stddev,
                                           self.mean + 2 * self.stddev)


class xavier_uniform(Initializer):
    def __init__(self, gain=1.0, seed=None):
        self.gain = gain

    def __call__(self, shape, dtype=torch.float32):
        fan_in, fan_out = _fans(tuple(shape))
        a = self.gain * math.sqrt(6.0 / (fan_in + fan_out))
        return torch.empty(tuple(shape), dtype=dtype).uniform_(-a, a)


class he_normal(Initializer):
    def __init__(self, a=0, mode="fan_in", nonlinearity="leaky_relu", seed=None):
        self.a = a

    def __call__(self, shape, dtype=torch.float32):
        fan_in, _ = _fans(tuple(shape))
        return torch.empty(tuple(shape), dtype=dtype).normal_(0.0, math.sqrt(2.0 / ((1 + self.a ** 2) * fan_in)))


_BY_NAME = {"zeros": zeros, "ones": ones, "constant": constant, "random_uniform": random_uniform,
            "random_normal": random_normal, "truncated_normal": truncated_normal, "xavier_uniform": xavier_uniform,
            "he_normal": he_normal}


def _resolve(init, default):
    """Accept an Initializer instance, its name, or any truthy flag (-> default)."""
    if isinstance(init, Initializer):
        return init
    if isinstance(init, str):
        return _BY_NAME[init]()
    if isinstance(init, type) and issubclass(init, Initializer):
        return init()
    return _BY_NAME[default]()
