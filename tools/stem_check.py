#!/usr/bin/env python
"""Row-ring stem kernel vs torch (bf16-rounded operands, fp32 math): conv(+BN+act)(+maxpool 3/2/1) cases.

    python tools/stem_check.py            # prints max-abs error per case, exits 1 on any failure
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch
import torch.nn.functional as F


def case(n, hw, cout, k, stride, pad, act, pool, seed=0, neg=False):
    from tlxcv_b200 import nn, runtime

    g = torch.Generator().manual_seed(seed)
    x = torch.randn(n, 3, hw, hw, generator=g)
    w = torch.randn(cout, 3, k, k, generator=g) * (2.0 / (3 * k * k)) ** 0.5
    gamma, beta = 0.75 + 0.5 * torch.rand(cout, generator=g), torch.randn(cout, generator=g) * 0.1
    if neg:
        beta = beta - 6.0
    mean, var = torch.randn(cout, generator=g) * 0.1, 0.75 + 0.5 * torch.rand(cout, generator=g)
    acts = {"relu": nn.ReLU, "relu6": nn.ReLU6, "leaky": lambda: nn.LeakyReLU(0.1), None: None}

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = nn.GroupConv2d(in_channels=3, out_channels=cout, kernel_size=k, stride=stride, padding=pad, b_init=None)
            self.bn = nn.BatchNorm2d(num_features=cout)
            self.a1 = acts[act]() if act else None
            self.pool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1) if pool else None

        def forward(self, x):
            y = self.bn(self.conv(x))
            y = self.a1(y) if self.a1 is not None else y
            return self.pool(y) if self.pool is not None else y

    net = Net()
    net.load_state_dict({"conv.filters": w, "bn.beta": beta, "bn.gamma": gamma, "bn.moving_mean": mean, "bn.moving_var": var})
    net = net.cuda().set_eval()
    plan, _, flat = runtime.get_plan(net, (x.cuda(),), {})
    out = plan.run(flat, graph=False)[0].cpu()
    kernels = [plan.op_info(i)["kernel"] for i in range(len(plan.spec.ops))]
    q = lambda t: t.bfloat16().float()
    y = F.conv2d(q(x), q(w), None, stride, pad)
    y = F.batch_norm(y, mean, var, gamma, beta, False, 0.0, 1e-5)
    y = {None: lambda t: t, "relu": F.relu, "relu6": F.relu6, "leaky": lambda t: F.leaky_relu(t, 0.1)}[act](y)
    y = q(y)
    if pool:
        y = F.max_pool2d(y, 3, 2, 1)
    err = float((out - y).abs().max())
    tol = 2.0 ** -7 * max(1.0, float(y.abs().max()))
    return err, tol, kernels, out.shape == y.shape


CASES = [
    dict(n=2, hw=64, cout=64, k=7, stride=2, pad=3, act="relu", pool=False),
    dict(n=2, hw=64, cout=64, k=7, stride=2, pad=3, act="relu", pool=True),
    dict(n=3, hw=224, cout=64, k=7, stride=2, pad=3, act="relu", pool=True),
    dict(n=2, hw=224, cout=64, k=7, stride=2, pad=3, act=None, pool=True, neg=True),
    dict(n=2, hw=30, cout=64, k=7, stride=2, pad=3, act="relu", pool=True),      # odd conv size 15 -> pooled 8
    dict(n=2, hw=32, cout=32, k=3, stride=2, pad=1, act="relu6", pool=False),
    dict(n=5, hw=224, cout=32, k=3, stride=2, pad=1, act="relu6", pool=False),
    dict(n=2, hw=24, cout=32, k=3, stride=1, pad=1, act="leaky", pool=False),
    dict(n=2, hw=224, cout=32, k=3, stride=1, pad=1, act="relu", pool=False),
    dict(n=1, hw=300, cout=32, k=3, stride=1, pad=1, act="leaky", pool=False),   # 150 pairs: two column tiles
    dict(n=2, hw=608, cout=32, k=3, stride=1, pad=1, act="leaky", pool=False),
    dict(n=150, hw=64, cout=64, k=7, stride=2, pad=3, act="relu", pool=True),    # more bands than SMs
]


def main():
    bad = 0
    for i, kw in enumerate(CASES):
        try:
            err, tol, kernels, shape_ok = case(seed=i, **kw)
            ok = shape_ok and err <= tol and any(k.startswith("stem_rowring") for k in kernels)
            print(f"{'ok  ' if ok else 'FAIL'} {kw} err={err:.4g} tol={tol:.4g} kernels={kernels}", flush=True)
        except Exception as e:  # noqa: BLE001
            ok = False
            print(f"FAIL {kw}: {type(e).__name__}: {e}", flush=True)
        bad += 0 if ok else 1
    sys.exit(1 if bad else 0)


if __name__ == "__main__":
    main()
