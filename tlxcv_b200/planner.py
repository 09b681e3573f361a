"""Graph -> fused plan specification (pure Python, runs without a GPU).

Fusion rule (the whole point of the B200 path): every ``conv`` absorbs the
``bn`` that follows it (folded to an fp32 per-channel scale/shift, applied in
the conv epilogue, never into bf16 weights), then an activation, then a
residual ``add`` whose other operand is already computed, then the activation
after the add:

    y = act2( act1( conv(x) * scale + shift ) + residual )

which covers ResNet ``out += identity; relu`` (resnet.py:152-155), ResNeXt
``tlx.add`` + ``tlx.relu`` (resnext.py:117-118), MobileNetV2 ``x + conv(x)``
(mobilenetv2.py:38) and DarkNet ``tlx.add(inputs, conv2)`` after LeakyReLU
(detection/backbones/darknet.py:155-159).  ``gap -> reshape -> linear`` becomes
GAP + a GEMM through the same tensor-core kernel with an fp32 epilogue.
"""
from __future__ import annotations

from dataclasses import dataclass, field
from typing import Any

from . import graph as _g

# op kinds / activations / dtypes: keep in sync with include/tlxcv_b200.h
(OP_IMPORT_NCHW, OP_CONV, OP_MAXPOOL, OP_GAP, OP_LINEAR, OP_ADD_ACT, OP_ARGMAX, OP_EXPORT_NCHW, OP_IMPORT_U8,
 OP_UPSAMPLE_CONCAT, OP_SOFTMAX, OP_SOFTMAX_CE, OP_AVGPOOL, OP_SPLAT_APPLY) = range(14)
ACT_NONE, ACT_RELU, ACT_RELU6, ACT_LEAKY = range(4)
DT_U8 = 4
DT_F32, DT_BF16, DT_I64, DT_ACT = 0, 1, 2, 3          # DT_ACT: bf16 in the default mode, f32 in validation mode
ROLE_INTERNAL, ROLE_INPUT, ROLE_OUTPUT = 0, 1, 2

_ACT = {None: ACT_NONE, "relu": ACT_RELU, "relu6": ACT_RELU6, "leaky": ACT_LEAKY}
OP_NAMES = ["import_nchw", "conv", "maxpool", "gap", "linear", "add_act", "argmax", "export_nchw", "import_u8_nhwc",
            "upsample_concat", "softmax", "softmax_ce", "avgpool", "splat_apply"]


class DerivedConv:
    """A conv whose packed filters are a FUNCTION of a module's filters, evaluated when the plan is built:

    * ``kind="dense"``: a grouped conv with C_in != C_out (ResNeSt's radix-major 3x3, groups = cardinality * radix,
      classification/resnest.py:103-112) as the dense conv with block-diagonal filters (zeros contribute exactly nothing);
    * ``kind="tile"``: a 1x1 conv applied to the SUM of ``n`` channel groups of its input (``add_n(split(x, radix))`` ->
      GAP -> conv, resnest.py:148-155) as one conv over all groups with the filters repeated along C_in.

    The cache fingerprint follows the source module's parameters (``_parameters`` is the same dict)."""

    def __init__(self, src, kind, n):
        self.src, self.kind, self.n = src, kind, n
        self._parameters, self._buffers = src._parameters, {}

    @property
    def filters(self):
        w = self.src.filters.detach()
        if self.kind == "tile":
            return w.repeat(1, self.n, 1, 1).contiguous()
        g = self.n
        kout, cg, r, s = w.shape
        dense = w.new_zeros((kout, cg * g, r, s))
        per = kout // g
        for i in range(g):
            dense[i * per:(i + 1) * per, i * cg:(i + 1) * cg] = w[i * per:(i + 1) * per]
        return dense

    @property
    def biases(self):
        return self.src.biases


@dataclass
class TensorSpec:
    n: int
    h: int
    w: int
    c: int            # logical channels (features for (N, F) tensors)
    dtype: int
    role: int = ROLE_INTERNAL


@dataclass
class OpSpec:
    kind: int
    in0: int
    out: int
    in1: int = -1
    r: int = 1
    s: int = 1
    stride: int = 1
    pad: int = 0
    dil: int = 1
    groups: int = 1
    act1: int = ACT_NONE
    alpha1: float = 0.0
    act2: int = ACT_NONE
    alpha2: float = 0.0
    conv: Any = None      # GroupConv2d / Linear module (parameters are read at plan build)
    norm: Any = None      # vision.NormalizeToTensor module (mean / std buffers)
    bn: Any = None        # BatchNorm module
    path: str = ""


@dataclass
class PlanSpec:
    tensors: list[TensorSpec] = field(default_factory=list)
    ops: list[OpSpec] = field(default_factory=list)
    inputs: list[int] = field(default_factory=list)       # plan tensor index per graph input (external NCHW f32)
    outputs: list[int] = field(default_factory=list)      # plan tensor index per graph output
    out_shapes: list[tuple] = field(default_factory=list)  # logical torch shapes of the outputs
    out_dtypes: list[int] = field(default_factory=list)

    def modules(self):
        """Parameter-holding modules in op order (for cache fingerprints)."""
        for op in self.ops:
            if op.conv is not None:
                yield op.conv
            if op.bn is not None:
                yield op.bn
            if op.norm is not None:
                yield op.norm

    def summary(self):
        return [(OP_NAMES[o.kind], o.path) for o in self.ops]


def _square(v, what, path):
    if v[0] != v[1]:
        raise NotImplementedError(f"{path}: non-square {what} {v} is not on the hot path")
    return int(v[0])


def lower(graph: _g.Graph) -> PlanSpec:
    """Fuse and lower a traced graph."""
    spec = PlanSpec()
    nodes = graph.nodes
    consumers: dict[int, list[int]] = {}
    for i, nd in enumerate(nodes):
        for t in nd.inputs:
            consumers.setdefault(t, []).append(i)
    graph_outputs = set(graph.outputs)

    gid2plan: dict[int, int] = {}          # graph tensor id -> plan tensor index
    produced_at: dict[int, int] = {}        # graph tensor id -> node index that made it available (-1: input)
    absorbed: set[int] = set()

    def new_tensor(shape, dtype, role=ROLE_INTERNAL):
        if len(shape) == 4:
            n, c, h, w = shape
        elif len(shape) == 2:
            (n, c), h, w = shape, 1, 1
        elif len(shape) == 1:
            n, c, h, w = shape[0], 1, 1, 1
        elif len(shape) == 0:
            n, c, h, w = 1, 1, 1, 1
        else:
            raise NotImplementedError(f"rank-{len(shape)} tensor on the plan")
        spec.tensors.append(TensorSpec(n, h, w, c, dtype, role))
        return len(spec.tensors) - 1

    def sole_consumer(tid, want_op):
        """Index of the only consumer of ``tid`` if it is a ``want_op`` node and ``tid`` is not a graph output."""
        cs = consumers.get(tid, [])
        if len(cs) == 1 and tid not in graph_outputs and nodes[cs[0]].op == want_op and cs[0] not in absorbed:
            return cs[0]
        return None

    # graph inputs: external NCHW fp32 -> internal NHWC (import op)
    for tid in graph.inputs:
        shape = graph.shapes[tid]
        if graph.dtypes[tid] in ("f32", "i64"):
            # (N, classes) fp32 logits or (N,) int64 labels handed to a loss / softmax head: used where they lie
            if (graph.dtypes[tid], len(shape)) not in (("f32", 2), ("i64", 1)):
                raise NotImplementedError(f"plan input {shape} {graph.dtypes[tid]}: only (N, K) fp32 and (N,) int64 besides images")
            ext = new_tensor(shape, DT_F32 if graph.dtypes[tid] == "f32" else DT_I64, ROLE_INPUT)
            spec.inputs.append(ext)
            gid2plan[tid] = ext
            produced_at[tid] = -1
            continue
        if len(shape) != 4:
            raise NotImplementedError(f"plan input must be (N, C, H, W), got {shape}")
        if graph.dtypes[tid] == "u8":
            # uint8 NHWC image batch: stays external until its (only) consumer, a vision.NormalizeToTensor node
            n, h, w, c = shape
            spec.tensors.append(TensorSpec(n, h, w, c, DT_U8, ROLE_INPUT))
            spec.inputs.append(len(spec.tensors) - 1)
            gid2plan[tid] = len(spec.tensors) - 1
            produced_at[tid] = -1
            continue
        ext = new_tensor(shape, DT_F32, ROLE_INPUT)
        spec.inputs.append(ext)
        inner = new_tensor(shape, DT_ACT)
        spec.ops.append(OpSpec(OP_IMPORT_NCHW, ext, inner, path="<input>"))
        gid2plan[tid] = inner
        produced_at[tid] = -1

    for i, nd in enumerate(nodes):
        if i in absorbed:
            continue
        ins = [gid2plan[t] for t in nd.inputs]
        if nd.op == "conv":
            a = nd.attrs
            op = OpSpec(OP_CONV, ins[0], -1, r=a["r"], s=a["s"], stride=_square(a["stride"], "stride", nd.path),
                        pad=_square(a["pad"], "padding", nd.path), dil=_square(a["dil"], "dilation", nd.path),
                        groups=a["groups"], conv=nd.module, path=nd.path)
            kout, cg = nd.module.filters.shape[:2]
            if a.get("dup_in", 1) > 1:
                op.conv = DerivedConv(nd.module, "tile", a["dup_in"])
            elif op.groups > 1 and not (kout == cg * op.groups and (cg == 1 or (64 % cg == 0 and kout % 64 == 0))):
                # only depthwise and C_in == C_out groups of 2..64 channels have grouped kernels: anything else runs dense
                op.conv, op.groups = DerivedConv(nd.module, "dense", op.groups), 1
            cur, last = nd.out, i
            j = sole_consumer(cur, "bn")
            if j is not None:
                op.bn = nodes[j].module
                absorbed.add(j)
                cur, last = nodes[j].out, j
            j = sole_consumer(cur, "act")
            if j is not None:
                op.act1, op.alpha1 = _ACT[nodes[j].attrs["kind"]], nodes[j].attrs["alpha"]
                absorbed.add(j)
                cur, last = nodes[j].out, j
            j = sole_consumer(cur, "add")
            if j is not None:
                other = [t for t in nodes[j].inputs if t != cur]
                # the other addend must already be available when this conv runs
                if len(other) == 1 and other[0] in produced_at and produced_at[other[0]] < i:
                    op.in1 = gid2plan[other[0]]
                    absorbed.add(j)
                    cur, last = nodes[j].out, j
                    k = sole_consumer(cur, "act")
                    if k is not None:
                        op.act2, op.alpha2 = _ACT[nodes[k].attrs["kind"]], nodes[k].attrs["alpha"]
                        absorbed.add(k)
                        cur, last = nodes[k].out, k
            op.out = new_tensor(graph.shapes[cur], DT_ACT)
            spec.ops.append(op)
            gid2plan[cur] = op.out
            produced_at[cur] = i
            continue
        if nd.op == "resize_u8":
            # only as the front of Resize -> Normalize -> ToTensor: the resize runs inside the import pass of the next node
            j = sole_consumer(nd.out, "normalize_u8")
            if j is None or spec.tensors[ins[0]].role != ROLE_INPUT:
                raise NotImplementedError(f"{nd.path}: Resize is supported on the uint8 input batch, directly in front of "
                                          "NormalizeToTensor")
            gid2plan[nd.out] = ins[0]
            produced_at[nd.out] = i
            continue
        if nd.op == "normalize_u8":
            src = spec.tensors[ins[0]]
            if src.dtype != DT_U8 or src.role != ROLE_INPUT:
                raise NotImplementedError(f"{nd.path}: NormalizeToTensor takes the uint8 NHWC input batch of the forward")
            out = new_tensor(graph.shapes[nd.out], DT_ACT)
            spec.ops.append(OpSpec(OP_IMPORT_U8, ins[0], out, conv=None, path=nd.path, norm=nd.module))
            gid2plan[nd.out] = out
            produced_at[nd.out] = i
            continue
        if nd.op in ("upsample", "concat"):
            # out[..., :C0] = in0 up-sampled r times, out[..., C0:] = in1 up-sampled s times (nearest): ONE pass writes the
            # concatenated map.  `concat([upsample(route), x])` of YOLOv3FPN (detection/yolov3.py:244,252-253) is one op.
            if nd.op == "upsample":
                op = OpSpec(OP_UPSAMPLE_CONCAT, ins[0], -1, r=nd.attrs["scale"], s=1, path=nd.path)
                cur = nd.out
                j = sole_consumer(cur, "concat")
                if j is not None and len(nodes[j].inputs) == 2:
                    other = [t for t in nodes[j].inputs if t != cur]
                    if len(other) == 1 and other[0] in produced_at and produced_at[other[0]] < i:
                        if nodes[j].inputs[0] == cur:
                            op.in1 = gid2plan[other[0]]
                        else:       # the up-sampled map is the second part
                            op.in0, op.in1, op.r, op.s = gid2plan[other[0]], ins[0], 1, nd.attrs["scale"]
                        absorbed.add(j)
                        cur = nodes[j].out
                op.out = new_tensor(graph.shapes[cur], DT_ACT)
                spec.ops.append(op)
                gid2plan[cur] = op.out
                produced_at[cur] = i
                continue
            acc = ins[0]
            for k, nxt in enumerate(ins[1:]):       # more than two parts: pairwise, left to right
                a, b = spec.tensors[acc], spec.tensors[nxt]
                last = k == len(ins) - 2
                out = new_tensor(graph.shapes[nd.out] if last else (a.n, a.c + b.c, a.h, a.w), DT_ACT)
                spec.ops.append(OpSpec(OP_UPSAMPLE_CONCAT, acc, out, in1=nxt, r=1, s=1, path=nd.path))
                acc = out
            gid2plan[nd.out] = acc
            produced_at[nd.out] = i
            continue
        if nd.op == "bn":
            raise NotImplementedError(f"{nd.path}: BatchNorm that does not follow a convolution is not on the hot path")
        if nd.op == "act":
            out = new_tensor(graph.shapes[nd.out], DT_ACT)
            spec.ops.append(OpSpec(OP_ADD_ACT, ins[0], out, act2=_ACT[nd.attrs["kind"]], alpha2=nd.attrs["alpha"],
                                   path=nd.path))
        elif nd.op == "add":
            op = OpSpec(OP_ADD_ACT, ins[0], -1, in1=ins[1], path=nd.path)
            cur = nd.out
            j = sole_consumer(cur, "act")
            if j is not None:
                op.act2, op.alpha2 = _ACT[nodes[j].attrs["kind"]], nodes[j].attrs["alpha"]
                absorbed.add(j)
                cur = nodes[j].out
            op.out = out = new_tensor(graph.shapes[cur], DT_ACT)
            spec.ops.append(op)
            gid2plan[cur] = out
            produced_at[cur] = i
            continue
        elif nd.op == "maxpool":
            a = nd.attrs
            out = new_tensor(graph.shapes[nd.out], DT_ACT)
            spec.ops.append(OpSpec(OP_MAXPOOL, ins[0], out, r=a["k"][0], s=a["k"][1],
                                   stride=_square(a["stride"], "stride", nd.path),
                                   pad=_square(a["pad"], "padding", nd.path), path=nd.path))
        elif nd.op == "avgpool":
            a = nd.attrs
            out = new_tensor(graph.shapes[nd.out], DT_ACT)
            spec.ops.append(OpSpec(OP_AVGPOOL, ins[0], out, r=a["k"][0], s=a["k"][1], stride=a["stride"][0],
                                   pad=a.get("pad", (0, 0))[0], path=nd.path))
        elif nd.op == "splat_apply":
            out = new_tensor(graph.shapes[nd.out], DT_ACT)
            spec.ops.append(OpSpec(OP_SPLAT_APPLY, ins[0], out, in1=ins[1], r=nd.attrs["radix"], groups=nd.attrs["cardinality"],
                                   path=nd.path))
        elif nd.op == "gap":
            out = new_tensor(graph.shapes[nd.out], DT_ACT)
            spec.ops.append(OpSpec(OP_GAP, ins[0], out, path=nd.path))
        elif nd.op == "reshape":
            src, dst = graph.shapes[nd.inputs[0]], graph.shapes[nd.out]
            n = src[0]
            flat_ok = (len(src) == 4 and src[2] == 1 and src[3] == 1) or len(src) == 2
            if not (flat_ok and len(dst) == 2 and dst[0] == n):
                raise NotImplementedError(
                    f"reshape {src} -> {dst}: only flattening a globally pooled (N, C, 1, 1) map is on the hot path "
                    "(NHWC storage makes other NCHW-order reshapes a data movement)")
            gid2plan[nd.out] = ins[0]              # pure alias in NHWC storage
            produced_at[nd.out] = produced_at[nd.inputs[0]]
            continue
        elif nd.op == "linear":
            out = new_tensor(graph.shapes[nd.out], DT_F32)
            spec.ops.append(OpSpec(OP_LINEAR, ins[0], out, conv=nd.module, path=nd.path))
        elif nd.op == "argmax":
            out = new_tensor(graph.shapes[nd.out], DT_I64)
            spec.ops.append(OpSpec(OP_ARGMAX, ins[0], out, path=nd.path))
        elif nd.op == "softmax":
            out = new_tensor(graph.shapes[nd.out], DT_F32)
            spec.ops.append(OpSpec(OP_SOFTMAX, ins[0], out, path=nd.path))
        elif nd.op == "softmax_ce":
            out = new_tensor(graph.shapes[nd.out], DT_F32)
            spec.ops.append(OpSpec(OP_SOFTMAX_CE, ins[0], out, in1=ins[1], path=nd.path))
        else:
            raise NotImplementedError(nd.op)
        gid2plan[nd.out] = out
        produced_at[nd.out] = i

    # graph outputs
    for tid in graph.outputs:
        pt = gid2plan[tid]
        t = spec.tensors[pt]
        shape = graph.shapes[tid]
        if t.role == ROLE_INPUT:
            raise NotImplementedError("a plan input returned unchanged")
        if t.dtype == DT_ACT:
            ext = new_tensor((t.n, t.c, t.h, t.w), DT_F32, ROLE_OUTPUT)
            spec.ops.append(OpSpec(OP_EXPORT_NCHW, pt, ext, path="<output>"))
            pt = ext
        else:
            if t.role == ROLE_OUTPUT:
                raise NotImplementedError("the same tensor returned twice")
            t.role = ROLE_OUTPUT
        spec.outputs.append(pt)
        spec.out_shapes.append(tuple(shape))
        spec.out_dtypes.append(spec.tensors[pt].dtype)
    return spec


# --------------------------------------------------------------------------- #
# tracing a module
# --------------------------------------------------------------------------- #
def _map_structure(obj, fn):
    if isinstance(obj, dict):
        return {k: _map_structure(v, fn) for k, v in obj.items()}
    if isinstance(obj, (list, tuple)):
        return type(obj)(_map_structure(v, fn) for v in obj)
    return fn(obj)


def trace(module, args, kwargs, is_tensor, shape_of):
    """Run ``module.forward`` on placeholders.

    Returns ``(graph, flat_inputs, out_structure)`` where ``flat_inputs`` are
    the real tensors in graph-input order and ``out_structure`` mirrors the
    forward's return value with graph-output positions at the leaves.
    """
    flat_inputs = []
    with _g.tracing() as g:
        g.names = {id(m): n for n, m in module.named_modules()}

        def to_sym(v):
            if is_tensor(v):
                flat_inputs.append(v)
                dt = str(getattr(v, "dtype", getattr(v, "kind", "")))
                shape = shape_of(v)
                if dt == "torch.uint8" or getattr(v, "u8", False):
                    return g.add_input(shape, "u8")
                if dt in ("torch.int64", "i64"):
                    return g.add_input(shape, "i64")
                if len(shape) == 2:
                    return g.add_input(shape, "f32")
                return g.add_input(shape, "act")
            return v

        s_args = _map_structure(list(args), to_sym)
        s_kwargs = _map_structure(dict(kwargs), to_sym)
        result = module.forward(*s_args, **s_kwargs)

        def to_slot(v):
            if isinstance(v, _g.SymTensor):
                g.outputs.append(v.id)
                return _OutSlot(len(g.outputs) - 1)
            return v

        structure = _map_structure(result, to_slot)
    if not g.outputs:
        raise ValueError("forward returned no tensors")
    return g, flat_inputs, structure


class _OutSlot:
    __slots__ = ("index",)

    def __init__(self, index):
        self.index = index


def fill_structure(structure, outputs):
    return _map_structure(structure, lambda v: outputs[v.index] if isinstance(v, _OutSlot) else v)


class Shape:
    """Stand-in for a real input tensor when planning on the CPU."""

    def __init__(self, *dims, u8=False, kind=""):
        self.dims = tuple(int(d) for d in dims)
        self.u8 = bool(u8)      # a uint8 (N, H, W, C) image batch (vision.NormalizeToTensor input)
        self.kind = kind        # "i64": (N,) int64 class labels


def plan_for_shapes(module, *args, **kwargs):
    """CPU-side planning from input shapes only (tests, tooling): pass ``Shape(N, C, H, W)``
    wherever the forward takes a tensor (also inside dicts, for DarkNet's ``{"images": ...}``)."""
    g, _, structure = trace(module, args, kwargs, lambda v: isinstance(v, Shape), lambda v: v.dims)
    return lower(g), structure
