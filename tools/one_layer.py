#!/usr/bin/env python
"""Run ONE fused conv layer repeatedly (for ncu captures and quick timing).

    python tools/one_layer.py --cin 64 --cout 256 --hw 56 --n 256 --k 1 [--stride 1] [--res] [--act relu] [--iters 5]
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    import tlxcv_b200 as tlx  # noqa: F401
    from tlxcv_b200 import nn, runtime

    ap = argparse.ArgumentParser()
    ap.add_argument("--cin", type=int, default=64)
    ap.add_argument("--cout", type=int, default=256)
    ap.add_argument("--hw", type=int, default=56)
    ap.add_argument("--n", type=int, default=256)
    ap.add_argument("--k", type=int, default=1)
    ap.add_argument("--stride", type=int, default=1)
    ap.add_argument("--groups", type=int, default=1)
    ap.add_argument("--res", action="store_true")
    ap.add_argument("--act", default="relu")
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--act-first", action="store_true", help="y = act(bn(conv(x))) + r (DarkNet) instead of act(bn(conv(x)) + r)")
    a = ap.parse_args()
    acts = {"none": None, "relu": nn.ReLU, "relu6": nn.ReLU6, "leaky": lambda: nn.LeakyReLU(0.1)}

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.pre = nn.GroupConv2d(in_channels=8, out_channels=a.cin, kernel_size=1, padding=0, b_init=None)
            self.conv = nn.GroupConv2d(in_channels=a.cin, out_channels=a.cout, kernel_size=a.k, stride=a.stride,
                                       padding=(a.k - 1) // 2, n_group=a.groups, b_init=None)
            self.bn = nn.BatchNorm2d(num_features=a.cout)
            self.act = acts[a.act]() if acts[a.act] else None

        def forward(self, x, r=None):
            y = self.bn(self.conv(self.pre(x)))
            if a.act_first:
                y = self.act(y) if self.act is not None else y
                return y + r if r is not None else y
            if r is not None:
                y = y + r
            return self.act(y) if self.act is not None else y

    net = Net().cuda().set_eval()
    x = torch.randn(a.n, 8, a.hw, a.hw, device="cuda")
    po = (a.hw + 2 * ((a.k - 1) // 2) - a.k) // a.stride + 1
    args = (x, torch.randn(a.n, a.cout, po, po, device="cuda")) if a.res else (x,)
    plan, _, flat = runtime.get_plan(net, args, {})
    outs = plan.alloc_outputs()
    for _ in range(a.iters):
        plan.run(flat, outs, graph=False)
    torch.cuda.synchronize()
    prof = plan.profile(flat, outs)
    for p in prof:
        if p["path"] == "conv":
            t = p["ms"] * 1e-3
            print(f"{p['kernel']} grid {p['grid']} {p['ms'] * 1e3:.1f} us  {p['flops'] / t / 1e12:.1f} TFLOP/s  "
                  f"{p['bytes'] / t / 1e9:.0f} GB/s (algorithmic)")


if __name__ == "__main__":
    main()
