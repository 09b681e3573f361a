"""Symbolic trace of a layer-API forward into a small op graph.

A ``tlxcv_b200.nn.Module`` called with real CUDA tensors is traced ONCE per
(input shapes, precision): its ``forward`` runs on ``SymTensor`` placeholders
and every layer / functional records a node here instead of computing.  The
planner (planner.py) then fuses conv+BN+act(+residual add+act) chains and lowers
the graph to the C-ABI plan that runs on the B200.  This replaces the
reference's 175 eager op launches per ResNet-50 forward with
one call across the boundary per forward.

Shapes are kept in the reference's logical NCHW order (``(N, C, H, W)`` or
``(N, F)``); the physical layout on the device is NHWC and is the planner's
business.
"""
from __future__ import annotations

import threading
from dataclasses import dataclass, field
from typing import Any

_state = threading.local()


def active() -> "Graph | None":
    return getattr(_state, "graph", None)


@dataclass
class Node:
    op: str                       # conv | bn | act | add | maxpool | gap | reshape | linear | argmax | upsample | concat | ...
    inputs: list[int]
    out: int
    attrs: dict[str, Any] = field(default_factory=dict)
    module: Any = None            # the layer holding the parameters (conv / bn / linear)
    path: str = ""                # module path, for error messages and per-layer reports


class SymTensor:
    """Placeholder for an activation during tracing."""

    __slots__ = ("graph", "id", "shape", "dtype", "stop_gradient")

    def __init__(self, graph, tid, shape, dtype="act"):
        self.graph, self.id, self.shape, self.dtype = graph, tid, tuple(int(s) for s in shape), dtype
        self.stop_gradient = False

    # the reference uses `x + y`, `out += identity` (resnet.py:154) and tlx.add
    def __add__(self, other):
        return self.graph.add(self, other)

    __radd__ = __add__
    __iadd__ = __add__

    @property
    def ndim(self):
        return len(self.shape)

    def size(self, dim=None):
        return self.shape if dim is None else self.shape[dim]

    def reshape(self, *shape):
        if len(shape) == 1 and isinstance(shape[0], (list, tuple)):
            shape = tuple(shape[0])
        return self.graph.reshape(self, shape)

    view = reshape

    def __repr__(self):
        return f"SymTensor(id={self.id}, shape={self.shape}, dtype={self.dtype})"

    # The reference's torch-backend helpers call torch functions on the activation itself
    # (detection/utils/ops.py:474-476: `F.interpolate(x, scale_factor=2.0)` behind `Interpolater`): route the few that
    # are on the hot path into the trace, refuse the rest loudly.
    @classmethod
    def __torch_function__(cls, func, types, args=(), kwargs=None):
        kwargs = kwargs or {}
        name = getattr(func, "__name__", str(func))
        if name == "interpolate":
            x = args[0] if args else kwargs["input"]
            return x.graph.interpolate(x, size=kwargs.get("size", args[1] if len(args) > 1 else None),
                                       scale_factor=kwargs.get("scale_factor", args[2] if len(args) > 2 else None),
                                       mode=kwargs.get("mode", args[3] if len(args) > 3 else "nearest"))
        if name in ("cat", "concat", "concatenate"):
            xs = args[0] if args else kwargs["tensors"]
            dim = kwargs.get("dim", kwargs.get("axis", args[1] if len(args) > 1 else 0))
            return xs[0].graph.concat(list(xs), dim)
        if name == "relu":
            return args[0].graph.act(args[0], "relu")
        raise NotImplementedError(f"torch.{name} on a traced activation is not on the B200 hot path")


class Graph:
    def __init__(self):
        self.nodes: list[Node] = []
        self.shapes: dict[int, tuple] = {}
        self.dtypes: dict[int, str] = {}
        self.inputs: list[int] = []
        self.outputs: list[int] = []
        self._path: list[str] = []

    # -- bookkeeping --------------------------------------------------------
    def new_tensor(self, shape, dtype="act") -> SymTensor:
        tid = len(self.shapes)
        t = SymTensor(self, tid, shape, dtype)
        self.shapes[tid] = t.shape
        self.dtypes[tid] = dtype
        return t

    def add_input(self, shape, dtype="act") -> SymTensor:
        t = self.new_tensor(shape, dtype)
        self.inputs.append(t.id)
        return t

    def _emit(self, op, ins, shape, attrs=None, module=None, dtype="act") -> SymTensor:
        for t in ins:
            if not isinstance(t, SymTensor) or t.graph is not self:
                raise TypeError(f"{op}: operands must be tensors traced in the same forward, got {type(t).__name__}")
        out = self.new_tensor(shape, dtype)
        self.nodes.append(Node(op, [t.id for t in ins], out.id, attrs or {}, module, ".".join(self._path)))
        return out

    # -- ops ----------------------------------------------------------------
    def conv(self, x, layer, dup_in=1) -> SymTensor:
        """``dup_in`` > 1: the input holds ``dup_in`` channel groups that the reference sums before the conv
        (``add_n(split(x, radix))`` in front of SplatConv's 1x1 conv): the planner tiles the filters along C_in instead."""
        n, c, h, w = _nchw(x, "GroupConv2d")
        kout, cg, r, s = layer.filters.shape
        g = layer.n_group
        if c != cg * g * dup_in:
            raise ValueError(f"GroupConv2d {'.'.join(self._path)}: input has {c} channels, filters expect {cg * g * dup_in}")
        if dup_in > 1 and (g != 1 or (r, s) != (1, 1)):
            raise NotImplementedError("summed channel groups in front of a conv: 1x1 dense convs only")
        (sh, sw), (ph, pw), (dh, dw) = layer.stride, layer.padding, layer.dilation
        p = (h + 2 * ph - dh * (r - 1) - 1) // sh + 1
        q = (w + 2 * pw - dw * (s - 1) - 1) // sw + 1
        attrs = dict(r=r, s=s, stride=(sh, sw), pad=(ph, pw), dil=(dh, dw), groups=g, dup_in=dup_in)
        return self._emit("conv", [x], (n, kout, p, q), attrs, layer)

    def splat_apply(self, x, logits, radix, cardinality) -> SymTensor:
        """Split attention of ResNeSt (classification/resnest.py:53-82,160-166): softmax over the radix axis of ``logits``
        ((N, radix*C, 1, 1), the conv's [cardinality][radix][C/cardinality] order), times the radix channel groups of
        ``x`` ((N, radix*C, H, W)), summed."""
        n, c, h, w = _nchw(x, "SplatConv")
        if tuple(logits.shape) != (n, c, 1, 1) or c % radix or (c // radix) % cardinality:
            raise ValueError(f"split attention: map {x.shape}, logits {logits.shape}, radix {radix}, cardinality {cardinality}")
        return self._emit("splat_apply", [x, logits], (n, c // radix, h, w), dict(radix=radix, cardinality=cardinality))

    def normalize_u8(self, x, layer) -> SymTensor:
        """uint8 NHWC image batch -> logical (N, C, H, W) activation, (x - mean) / std per channel."""
        if x.dtype != "u8" or len(x.shape) != 4:
            raise TypeError(f"NormalizeToTensor {'.'.join(self._path)}: expects the uint8 (N, H, W, C) image batch")
        n, h, w, c = x.shape
        if c != len(layer.mean) or c > 4:
            raise ValueError(f"NormalizeToTensor: {c} channels vs {len(layer.mean)} mean/std entries (at most 4)")
        return self._emit("normalize_u8", [x], (n, c, h, w), {}, layer)

    def resize_u8(self, x, size) -> SymTensor:
        """uint8 (N, Hs, Ws, C) image batch -> uint8 (N, H, W, C), OpenCV INTER_LINEAR (the reference's ``Resize``)."""
        if x.dtype != "u8" or len(x.shape) != 4:
            raise TypeError(f"Resize {'.'.join(self._path)}: expects the uint8 (N, H, W, C) image batch")
        n, _, _, c = x.shape
        return self._emit("resize_u8", [x], (n, int(size[0]), int(size[1]), c), {}, dtype="u8")

    def bn(self, x, layer) -> SymTensor:
        if x.shape[1] != layer.gamma.shape[0]:
            raise ValueError(f"BatchNorm {'.'.join(self._path)}: {x.shape[1]} channels vs {layer.gamma.shape[0]} features")
        return self._emit("bn", [x], x.shape, dict(eps=float(layer.epsilon)), layer)

    def act(self, x, kind, alpha=0.0) -> SymTensor:
        return self._emit("act", [x], x.shape, dict(kind=kind, alpha=float(alpha)))

    def add(self, a, b) -> SymTensor:
        if not isinstance(a, SymTensor) or not isinstance(b, SymTensor):
            raise TypeError("add: only tensor + tensor is on the hot path")
        if a.shape != b.shape:
            raise ValueError(f"add: shape mismatch {a.shape} vs {b.shape}")
        return self._emit("add", [a, b], a.shape)

    def interpolate(self, x, size=None, scale_factor=None, mode="nearest") -> SymTensor:
        """Nearest-neighbour up-sampling by an integer factor (YOLOv3FPN route, detection/yolov3.py:252-253)."""
        n, c, h, w = _nchw(x, "interpolate")
        if mode != "nearest":
            raise NotImplementedError(f"interpolate(mode={mode!r}): only nearest is on the hot path")
        if size is not None:
            sh, sw = (size, size) if isinstance(size, int) else tuple(size)
            if sh % h or sw % w or sh // h != sw // w:
                raise NotImplementedError(f"interpolate to {size} from {(h, w)}: only integer up-scaling")
            k = sh // h
        else:
            if scale_factor is None or float(scale_factor) != int(scale_factor) or int(scale_factor) < 1:
                raise NotImplementedError(f"interpolate(scale_factor={scale_factor}): only integer up-scaling")
            k = int(scale_factor)
        return self._emit("upsample", [x], (n, c, h * k, w * k), dict(scale=k))

    def concat(self, xs, axis=1) -> SymTensor:
        """Channel concatenation of (N, C_i, H, W) maps (``tlx.concat([route, x], axis=1)``, detection/yolov3.py:244)."""
        if len(xs) < 2:
            raise ValueError("concat needs at least two tensors")
        shapes = [_nchw(t, "concat") for t in xs]
        if axis not in (1, -3):
            raise NotImplementedError("concat: only along the channel axis of NCHW maps")
        if any((s[0], s[2], s[3]) != (shapes[0][0], shapes[0][2], shapes[0][3]) for s in shapes):
            raise ValueError(f"concat: shape mismatch {shapes}")
        n, _, h, w = shapes[0]
        return self._emit("concat", list(xs), (n, sum(s[1] for s in shapes), h, w))

    def maxpool(self, x, k, stride, pad) -> SymTensor:
        n, c, h, w = _nchw(x, "MaxPool2d")
        (kh, kw), (sh, sw), (ph, pw) = k, stride, pad
        p = (h + 2 * ph - kh) // sh + 1
        q = (w + 2 * pw - kw) // sw + 1
        return self._emit("maxpool", [x], (n, c, p, q), dict(k=(kh, kw), stride=(sh, sw), pad=(ph, pw)))

    def avgpool(self, x, k, stride, pad) -> SymTensor:
        n, c, h, w = _nchw(x, "AvgPool2d")
        (kh, kw), (sh, sw), (ph, pw) = k, stride, pad
        if ph != pw or kh != kw or sh != sw or 2 * ph > kh:
            raise NotImplementedError("AvgPool2d: only square windows with symmetric padding are on the hot path")
        if ph == 0 and ((h - kh) % sh or (w - kw) % sw):
            raise NotImplementedError(f"AvgPool2d({kh}, {sh}) on a {h}x{w} map: windows must tile the map exactly")
        if (kh, sh, ph) == (1, 1, 0):
            return x                # AvgPool2d(1, 1): ResNeSt's avg_down shortcut of a stride-1 block (classification/resnest.py:263-269)
        return self._emit("avgpool", [x], (n, c, (h + 2 * ph - kh) // sh + 1, (w + 2 * pw - kw) // sw + 1),
                          dict(k=(kh, kw), stride=(sh, sw), pad=(ph, pw)))

    def gap(self, x) -> SymTensor:
        n, c, h, w = _nchw(x, "AdaptiveAvgPool2d")
        return self._emit("gap", [x], (n, c, 1, 1))

    def reshape(self, x, shape) -> SymTensor:
        numel = 1
        for s in x.shape:
            numel *= s
        shape = list(int(s) for s in shape)
        if shape.count(-1) > 1:
            raise ValueError("reshape: at most one -1")
        if -1 in shape:
            known = 1
            for s in shape:
                if s != -1:
                    known *= s
            shape[shape.index(-1)] = numel // known
        prod = 1
        for s in shape:
            prod *= s
        if prod != numel:
            raise ValueError(f"reshape: cannot view {x.shape} as {tuple(shape)}")
        return self._emit("reshape", [x], tuple(shape))

    def linear(self, x, layer) -> SymTensor:
        if len(x.shape) != 2:
            raise ValueError(f"Linear expects (N, F) input, got {x.shape}")
        fin, fout = layer.weights.shape
        if x.shape[1] != fin:
            raise ValueError(f"Linear {'.'.join(self._path)}: input has {x.shape[1]} features, weights expect {fin}")
        return self._emit("linear", [x], (x.shape[0], fout), {}, layer, dtype="f32")

    def softmax(self, x, axis=-1) -> SymTensor:
        if len(x.shape) != 2 or axis not in (-1, 1) or x.dtype != "f32":
            raise NotImplementedError("softmax: only over the class axis of (N, classes) fp32 logits")
        return self._emit("softmax", [x], x.shape, {}, dtype="f32")

    def softmax_ce(self, logits, target) -> SymTensor:
        """Mean softmax cross-entropy of (N, classes) logits against (N,) int64 labels -> scalar."""
        if len(logits.shape) != 2 or logits.dtype != "f32":
            raise NotImplementedError("softmax_cross_entropy_with_logits: (N, classes) fp32 logits")
        if target.dtype != "i64" or tuple(target.shape) != (logits.shape[0],):
            raise NotImplementedError("softmax_cross_entropy_with_logits: target must be (N,) int64 class labels")
        return self._emit("softmax_ce", [logits, target], (), {}, dtype="f32")

    def argmax(self, x, axis=-1) -> SymTensor:
        if len(x.shape) != 2 or axis not in (-1, 1):
            raise NotImplementedError("argmax: only over the class axis of (N, classes) logits")
        return self._emit("argmax", [x], (x.shape[0],), {}, dtype="i64")


def _nchw(x, who):
    if not isinstance(x, SymTensor):
        raise TypeError(f"{who}: expected a traced tensor, got {type(x).__name__}")
    if len(x.shape) != 4:
        raise ValueError(f"{who}: expected (N, C, H, W) input, got {x.shape}")
    return x.shape


class tracing:
    """Context manager installing a fresh Graph as the active trace."""

    def __enter__(self) -> Graph:
        if active() is not None:
            raise RuntimeError("nested tracing")
        _state.graph = Graph()
        return _state.graph

    def __exit__(self, *exc):
        _state.graph = None
        return False
