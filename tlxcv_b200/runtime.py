"""ctypes binding of ``libtlxcv_b200.so`` and the per-module plan cache.

``run_module`` is what every ``tlxcv_b200.nn.Module.__call__`` ends in: it
traces the module (once per input signature), builds a C-ABI plan from the
fused graph and the module's parameter device pointers, and runs it on the
caller's current CUDA stream.  PyTorch is used only for device memory and
streams.  There is no fallback: without the extension, a B200 and CUDA inputs
this raises.
"""
from __future__ import annotations

import ctypes as C
import os
import threading

import torch

from . import planner

# TLXCV_B200_LIB: another build of the same library (A/B timing of kernel variants on one box); never a fallback
_LIB_PATH = os.environ.get("TLXCV_B200_LIB") or os.path.join(os.path.dirname(os.path.abspath(__file__)), "libtlxcv_b200.so")
ABI_VERSION = 1

PREC_BF16, PREC_F32 = 0, 1


class B200RuntimeError(RuntimeError):
    pass


class TensorDesc(C.Structure):
    _fields_ = [("n", C.c_int32), ("h", C.c_int32), ("w", C.c_int32), ("c", C.c_int32), ("dtype", C.c_int32),
                ("role", C.c_int32)]


class OpDesc(C.Structure):
    _fields_ = [("kind", C.c_int32), ("in0", C.c_int32), ("in1", C.c_int32), ("out", C.c_int32),
                ("r", C.c_int32), ("s", C.c_int32), ("stride", C.c_int32), ("pad", C.c_int32), ("dil", C.c_int32),
                ("groups", C.c_int32), ("act1", C.c_int32), ("alpha1", C.c_float), ("act2", C.c_int32),
                ("alpha2", C.c_float), ("filters", C.c_void_p), ("bias", C.c_void_p), ("bn_gamma", C.c_void_p),
                ("bn_beta", C.c_void_p), ("bn_mean", C.c_void_p), ("bn_var", C.c_void_p), ("bn_eps", C.c_float),
                ("reserved", C.c_int32)]


class OpInfo(C.Structure):
    _fields_ = [("kernel", C.c_char * 48), ("launches", C.c_int32), ("bound", C.c_int32), ("flops", C.c_double),
                ("bytes", C.c_double), ("grid", C.c_int32), ("block", C.c_int32), ("smem_bytes", C.c_int32),
                ("tile_n", C.c_int32)]


# every symbol include/tlxcv_b200.h declares (tests/test_abi.py checks the library exports them all)
EXPORTS = {
    "tlxcv_abi_version": (C.c_int, []),
    "tlxcv_create": (C.c_int, [C.c_int, C.POINTER(C.c_void_p)]),
    "tlxcv_destroy": (C.c_int, [C.c_void_p]),
    "tlxcv_last_error": (C.c_char_p, [C.c_void_p]),
    "tlxcv_device_sm_count": (C.c_int, [C.c_void_p]),
    "tlxcv_plan_build": (C.c_int, [C.c_void_p, C.POINTER(TensorDesc), C.c_int, C.POINTER(OpDesc), C.c_int, C.c_int,
                                   C.c_void_p, C.POINTER(C.c_void_p)]),
    "tlxcv_plan_destroy": (C.c_int, [C.c_void_p]),
    "tlxcv_plan_run": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_int]),
    "tlxcv_plan_run_host": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p, C.c_int]),
    "tlxcv_plan_profile": (C.c_int, [C.c_void_p, C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_void_p,
                                     C.POINTER(C.c_float), C.c_int]),
    "tlxcv_plan_num_ops": (C.c_int, [C.c_void_p]),
    "tlxcv_plan_num_launches": (C.c_int, [C.c_void_p]),
    "tlxcv_plan_op_info": (C.c_int, [C.c_void_p, C.c_int, C.POINTER(OpInfo)]),
    "tlxcv_plan_workspace_bytes": (C.c_size_t, [C.c_void_p]),
    "tlxcv_plan_read_tensor": (C.c_int, [C.c_void_p, C.c_int, C.c_void_p, C.c_size_t, C.c_void_p]),
}

_lib = None
_lib_lock = threading.Lock()
_contexts: dict[int, "Context"] = {}


def load_library():
    """Load the C-ABI library; fails loudly (no fallback) when it has not been built."""
    global _lib
    with _lib_lock:
        if _lib is None:
            if not os.path.isfile(_LIB_PATH):
                raise B200RuntimeError(
                    f"{_LIB_PATH} is missing: build it with `make -C tlxcv_b200/csrc` (or __graft_entry__.build()); "
                    "tlxcv_b200 has no CPU or eager fallback")
            lib = C.CDLL(_LIB_PATH)
            for name, (res, args) in EXPORTS.items():
                fn = getattr(lib, name)
                fn.restype, fn.argtypes = res, args
            if lib.tlxcv_abi_version() != ABI_VERSION:
                raise B200RuntimeError("libtlxcv_b200.so ABI version mismatch; rebuild")
            _lib = lib
    return _lib


class Context:
    def __init__(self, device: int):
        self.lib = load_library()
        self.device = device
        h = C.c_void_p()
        rc = self.lib.tlxcv_create(device, C.byref(h))
        if rc != 0:
            raise B200RuntimeError(f"tlxcv_create({device}) failed ({rc}): {self.lib.tlxcv_last_error(None).decode()}")
        self.handle = h

    def error(self) -> str:
        return self.lib.tlxcv_last_error(self.handle).decode()

    @property
    def sm_count(self) -> int:
        return self.lib.tlxcv_device_sm_count(self.handle)


def context(device: int) -> Context:
    if device not in _contexts:
        _contexts[device] = Context(device)
    return _contexts[device]


_TORCH_DTYPE = {planner.DT_F32: torch.float32, planner.DT_I64: torch.int64}


def _ptr(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _param(mod, name, device):
    p = getattr(mod, name, None)
    if p is None:
        return None
    if p.device != device or p.dtype != torch.float32 or not p.is_contiguous():
        raise B200RuntimeError(
            f"parameter {name!r} of {type(mod).__name__} is {p.dtype} on {p.device}; the B200 path needs contiguous "
            f"fp32 parameters on {device} (call model.to('{device}'))")
    return p.detach()


class Plan:
    """A built C-ABI plan plus what is needed to call it from torch tensors."""

    def __init__(self, spec: planner.PlanSpec, device: torch.device, precision: int):
        self.spec, self.device, self.precision = spec, device, precision
        self.ctx = context(device.index if device.index is not None else torch.cuda.current_device())
        lib = self.ctx.lib
        tens = (TensorDesc * len(spec.tensors))()
        for i, t in enumerate(spec.tensors):
            tens[i] = TensorDesc(t.n, t.h, t.w, t.c, t.dtype, t.role)
        ops = (OpDesc * len(spec.ops))()
        self._keep = []
        self._norm_keep = []   # mean / std vectors of import_u8 ops: read by the kernel at every run
        for i, o in enumerate(spec.ops):
            d = OpDesc(kind=o.kind, in0=o.in0, in1=o.in1, out=o.out, r=o.r, s=o.s, stride=o.stride, pad=o.pad,
                       dil=o.dil, groups=o.groups, act1=o.act1, alpha1=o.alpha1, act2=o.act2, alpha2=o.alpha2,
                       bn_eps=1e-5)
            if o.conv is not None:
                w = _param(o.conv, "filters" if o.kind == planner.OP_CONV else "weights", device)
                b = _param(o.conv, "biases", device)
                self._keep += [w, b]
                d.filters, d.bias = _ptr(w), _ptr(b)
            if getattr(o, "norm", None) is not None:
                mean = o.norm.mean.detach().to(device=device, dtype=torch.float32).contiguous().clone()
                std = o.norm.std.detach().to(device=device, dtype=torch.float32).contiguous().clone()
                self._norm_keep += [mean, std]
                d.bn_mean, d.bn_var = _ptr(mean), _ptr(std)
            if o.bn is not None:
                g, be = _param(o.bn, "gamma", device), _param(o.bn, "beta", device)
                mu, var = _param(o.bn, "moving_mean", device), _param(o.bn, "moving_var", device)
                self._keep += [g, be, mu, var]
                d.bn_gamma, d.bn_beta, d.bn_mean, d.bn_var = _ptr(g), _ptr(be), _ptr(mu), _ptr(var)
                d.bn_eps = float(o.bn.epsilon)
            ops[i] = d
        h = C.c_void_p()
        stream = torch.cuda.current_stream(device).cuda_stream
        with torch.cuda.device(device):
            rc = lib.tlxcv_plan_build(self.ctx.handle, tens, len(spec.tensors), ops, len(spec.ops), precision,
                                      C.c_void_p(stream), C.byref(h))
        if rc != 0:
            raise B200RuntimeError(f"tlxcv_plan_build failed ({rc}): {self.ctx.error()}")
        self.handle = h
        self.n_in, self.n_out = len(spec.inputs), len(spec.outputs)
        # the C ABI takes output pointers in tensor-table order; the caller's outputs are in forward-return order
        self._out_order = sorted(range(self.n_out), key=lambda i: spec.outputs[i])
        self.fingerprint = param_fingerprint(spec)
        self._keep = None      # the plan owns packed copies; the fp32 parameters are not referenced after build

    def __del__(self):
        h, self.handle = getattr(self, "handle", None), None
        if h is not None and _lib is not None:
            _lib.tlxcv_plan_destroy(h)

    # -- execution ------------------------------------------------------------
    def alloc_outputs(self):
        outs = []
        for shape, dt in zip(self.spec.out_shapes, self.spec.out_dtypes):
            outs.append(torch.empty(shape, dtype=_TORCH_DTYPE[dt], device=self.device))
        return outs

    def run(self, inputs, outputs=None, graph=True):
        if len(inputs) != self.n_in:
            raise ValueError(f"plan takes {self.n_in} inputs")
        for x, ti in zip(inputs, self.spec.inputs):
            t = self.spec.tensors[ti]
            if t.role == planner.ROLE_INPUT and t.dtype in (planner.DT_F32, planner.DT_I64) and t.h == 1 and t.w == 1 and \
                    (t.dtype == planner.DT_I64 or x.dim() == 2):
                want = (t.n,) if t.dtype == planner.DT_I64 else (t.n, t.c)
                dt = torch.int64 if t.dtype == planner.DT_I64 else torch.float32
                if tuple(x.shape) != want or x.dtype != dt or not x.is_contiguous() or x.device != self.device:
                    raise B200RuntimeError(f"plan input must be contiguous {dt} {want} on {self.device}, got {x.dtype} "
                                           f"{tuple(x.shape)} on {x.device}")
                continue
            if t.dtype == planner.DT_U8:
                if tuple(x.shape) != (t.n, t.h, t.w, t.c) or x.dtype != torch.uint8 or not x.is_contiguous() \
                        or x.device != self.device:
                    raise B200RuntimeError(f"plan input must be contiguous uint8 NHWC {(t.n, t.h, t.w, t.c)} on "
                                           f"{self.device}, got {x.dtype} {tuple(x.shape)} on {x.device}")
                continue
            if tuple(x.shape) != (t.n, t.c, t.h, t.w) or x.dtype != torch.float32 or not x.is_contiguous() \
                    or x.device != self.device:
                raise B200RuntimeError(f"plan input must be contiguous fp32 NCHW {(t.n, t.c, t.h, t.w)} on "
                                       f"{self.device}, got {x.dtype} {tuple(x.shape)} on {x.device}")
        outs = outputs if outputs is not None else self.alloc_outputs()
        ins = (C.c_void_p * self.n_in)(*[x.data_ptr() for x in inputs])
        ous = (C.c_void_p * self.n_out)(*[outs[i].data_ptr() for i in self._out_order])
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.ctx.lib.tlxcv_plan_run(self.handle, ins, ous, C.c_void_p(stream), 1 if graph else 0)
        if rc != 0:
            raise B200RuntimeError(f"tlxcv_plan_run failed ({rc}): {self.ctx.error()}")
        return outs

    def run_host(self, host_inputs, host_outputs, graph=True):
        """Pinned host tensors in, pinned host tensors out (H2D + run + D2H on the current stream)."""
        for t in list(host_inputs) + list(host_outputs):
            if t.device.type != "cpu" or not t.is_pinned() or not t.is_contiguous():
                raise B200RuntimeError("run_host needs contiguous pinned host tensors")
        ins = (C.c_void_p * self.n_in)(*[x.data_ptr() for x in host_inputs])
        ous = (C.c_void_p * self.n_out)(*[host_outputs[i].data_ptr() for i in self._out_order])
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.ctx.lib.tlxcv_plan_run_host(self.handle, ins, ous, C.c_void_p(stream), 1 if graph else 0)
        if rc != 0:
            raise B200RuntimeError(f"tlxcv_plan_run_host failed ({rc}): {self.ctx.error()}")

    def profile(self, inputs, outputs=None):
        """Per-op device milliseconds + the library's per-op report."""
        outs = outputs if outputs is not None else self.alloc_outputs()
        n = self.ctx.lib.tlxcv_plan_num_ops(self.handle)
        ms = (C.c_float * n)()
        ins = (C.c_void_p * self.n_in)(*[x.data_ptr() for x in inputs])
        ous = (C.c_void_p * self.n_out)(*[outs[i].data_ptr() for i in self._out_order])
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.ctx.lib.tlxcv_plan_profile(self.handle, ins, ous, C.c_void_p(stream), ms, n)
        if rc != 0:
            raise B200RuntimeError(f"tlxcv_plan_profile failed ({rc}): {self.ctx.error()}")
        return [dict(self.op_info(i), ms=ms[i]) for i in range(n)]

    def op_info(self, i):
        info = OpInfo()
        self.ctx.lib.tlxcv_plan_op_info(self.handle, i, C.byref(info))
        o = self.spec.ops[i]
        return dict(index=i, op=planner.OP_NAMES[o.kind], path=o.path, kernel=info.kernel.decode(),
                    launches=info.launches, bound="tensor" if info.bound else "hbm", flops=info.flops,
                    bytes=info.bytes, grid=info.grid, block=info.block, smem=info.smem_bytes, tile_n=info.tile_n)

    @property
    def num_launches(self):
        return self.ctx.lib.tlxcv_plan_num_launches(self.handle)

    @property
    def workspace_bytes(self):
        return self.ctx.lib.tlxcv_plan_workspace_bytes(self.handle)

    def read_tensor(self, index):
        """Debug: copy internal tensor ``index`` out as (N, H, W, C_storage) in the plan's activation dtype."""
        t = self.spec.tensors[index]
        dt = torch.float32 if self.precision == PREC_F32 else torch.bfloat16
        cs = 4 if t.c <= 4 else (t.c if self.precision == PREC_F32 else (t.c + 7) // 8 * 8)
        buf = torch.empty((t.n, t.h, t.w, cs), dtype=dt, device=self.device)
        stream = torch.cuda.current_stream(self.device).cuda_stream
        rc = self.ctx.lib.tlxcv_plan_read_tensor(self.handle, index, C.c_void_p(buf.data_ptr()),
                                                 buf.numel() * buf.element_size(), C.c_void_p(stream))
        if rc < 0:
            raise B200RuntimeError(f"tlxcv_plan_read_tensor failed ({rc}): {self.ctx.error()}")
        return buf


def param_fingerprint(spec):
    """(data_ptr, version) of every parameter a plan packed: a changed weight rebuilds the plan."""
    fp = []
    for m in spec.modules():
        for p in list(m._parameters.values()) + list(m._buffers.values()):
            if p is not None:
                fp.append((p.data_ptr(), p._version))
    return tuple(fp)


def default_precision() -> int:
    """``TLXCV_B200_VALIDATE_FP32=1`` selects the fp32 validation mode."""
    return PREC_F32 if os.environ.get("TLXCV_B200_VALIDATE_FP32", "0") not in ("", "0") else PREC_BF16


def get_plan(module, args, kwargs, precision=None):
    """Trace + build (or fetch from the module's cache) the plan for these inputs."""
    precision = default_precision() if precision is None else precision

    def is_tensor(v):
        return isinstance(v, torch.Tensor)

    probe = []
    planner._map_structure([list(args), dict(kwargs)], lambda v: probe.append(v) if is_tensor(v) else None)
    if not probe:
        raise B200RuntimeError("module called without tensor inputs")
    device = probe[0].device
    if device.type != "cuda":
        raise B200RuntimeError(
            f"tlxcv_b200 executes on a B200 only; got an input on {device} (there is no CPU fallback — move the "
            "model and the inputs to 'cuda')")
    key = (tuple((tuple(t.shape), str(t.dtype)) for t in probe), str(device), precision)
    cache = module.__dict__.setdefault("_b200_plans", {})
    entry = cache.get(key)
    if entry is not None:
        plan, structure = entry
        if plan.fingerprint == param_fingerprint(plan.spec):
            return plan, structure, probe
        del cache[key]
    graph, flat_inputs, structure = planner.trace(module, args, kwargs, is_tensor, lambda t: tuple(t.shape))
    spec = planner.lower(graph)
    plan = Plan(spec, device, precision)
    cache[key] = (plan, structure)
    return plan, structure, flat_inputs


def run_module(module, args, kwargs):
    plan, structure, flat_inputs = get_plan(module, args, kwargs)
    ins = [x if (x.dtype in (torch.float32, torch.uint8, torch.int64) and x.is_contiguous()) else
           (x.contiguous() if x.dtype in (torch.uint8, torch.int64) else x.float().contiguous()) for x in flat_inputs]
    outs = plan.run(ins)
    return planner.fill_structure(structure, outs)
