#!/usr/bin/env python
"""Time ONE bottleneck tail (3x3 conv + BN + ReLU -> 1x1 conv + BN + residual + ReLU) as the chain kernel runs it.

    python tools/chain_layer.py --cin 64 --mid 64 --cout 256 --hw 56 --n 256
"""
from __future__ import annotations

import argparse
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from tlxcv_b200 import nn, runtime

    ap = argparse.ArgumentParser()
    ap.add_argument("--cin", type=int, default=64)
    ap.add_argument("--mid", type=int, default=64)
    ap.add_argument("--cout", type=int, default=256)
    ap.add_argument("--hw", type=int, default=56)
    ap.add_argument("--n", type=int, default=256)
    a = ap.parse_args()

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.pre = nn.GroupConv2d(in_channels=8, out_channels=a.cin, kernel_size=1, padding=0, b_init=None)
            self.conv2 = nn.GroupConv2d(in_channels=a.cin, out_channels=a.mid, kernel_size=3, padding=1, b_init=None)
            self.bn2 = nn.BatchNorm2d(num_features=a.mid)
            self.conv3 = nn.GroupConv2d(in_channels=a.mid, out_channels=a.cout, kernel_size=1, padding=0, b_init=None)
            self.bn3 = nn.BatchNorm2d(num_features=a.cout)
            self.relu = nn.ReLU()

        def forward(self, x, r):
            out = self.bn3(self.conv3(self.relu(self.bn2(self.conv2(self.pre(x))))))
            out += r
            return self.relu(out)

    net = Net().cuda().set_eval()
    x = torch.randn(a.n, 8, a.hw, a.hw, device="cuda")
    r = torch.randn(a.n, a.cout, a.hw, a.hw, device="cuda")
    plan, _, flat = runtime.get_plan(net, (x, r), {})
    outs = plan.alloc_outputs()
    for _ in range(3):
        plan.run(flat, outs, graph=False)
    torch.cuda.synchronize()
    best = {}
    for _ in range(3):
        for p in plan.profile(flat, outs):
            if p["kernel"].startswith("conv_chain") or p["path"] in ("conv2", "conv3"):
                best[p["kernel"]] = min(best.get(p["kernel"], 1e9), p["ms"])
    print(" ".join(f"{k} {v * 1e3:.1f} us" for k, v in best.items()))


if __name__ == "__main__":
    main()
