#!/usr/bin/env python
"""Sweep the result-preserving tile switches over single conv layers: does a static rule of tc_conv_prepare leave time on the table?

    python tools/tune_sweep.py [layer ...]

For every layer (ResNet-50 bs256 shapes that run on the plain conv_tcgen05 kernel) and every combination of
TLXCV_FORCE_BLOCK_N {auto, 64, 128, 256} x TLXCV_DEBUG_2SM {auto, 0, 1} x TLXCV_DEBUG_RING {auto, 2, 4} prints the time of the conv launch
(best of 3 profiling passes, microseconds), the automatic choice first and then every combination that beats it by > 2 %.
"""
from __future__ import annotations

import itertools
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LAYERS = [  # name, cin, cout, hw, k, stride, residual
    ("l1.0.conv1", 64, 64, 56, 1, 1, False),
    ("l1.x.conv1", 256, 64, 56, 1, 1, False),
    ("l2.0.conv1", 256, 128, 56, 1, 1, False),
    ("l2.0.conv2", 128, 128, 56, 3, 2, False),
    ("l2.x.conv1", 512, 128, 28, 1, 1, False),
    ("l3.0.conv1", 512, 256, 28, 1, 1, False),
    ("l3.0.conv2", 256, 256, 28, 3, 2, False),
    ("l3.0.conv3", 256, 1024, 14, 1, 1, False),
    ("l3.0.down", 512, 1024, 28, 1, 2, False),
    ("l3.x.conv1", 1024, 256, 14, 1, 1, False),
    ("l3.x.conv2", 256, 256, 14, 3, 1, False),
    ("l3.x.conv3", 256, 1024, 14, 1, 1, True),
    ("l4.0.conv1", 1024, 512, 14, 1, 1, False),
    ("l4.0.conv2", 512, 512, 14, 3, 2, False),
    ("l4.0.conv3", 512, 2048, 7, 1, 1, False),
    ("l4.0.down", 1024, 2048, 14, 1, 2, False),
    ("l4.x.conv1", 2048, 512, 7, 1, 1, False),
    ("l4.x.conv2", 512, 512, 7, 3, 1, False),
    ("l4.x.conv3", 512, 2048, 7, 1, 1, True),
]
ENV = {"TLXCV_FORCE_BLOCK_N": [None, "64", "128", "256"], "TLXCV_DEBUG_2SM": [None, "0", "1"], "TLXCV_DEBUG_RING": [None, "2", "4"]}


def main():
    import torch

    import tlxcv_b200 as tlx  # noqa: F401
    from tlxcv_b200 import nn, runtime

    os.environ["TLXCV_NO_CHAIN"] = "1"
    os.environ["TLXCV_NO_DUAL"] = "1"
    os.environ["TLXCV_NO_SLAB"] = "1"
    n = 256
    only = sys.argv[1:]
    for name, cin, cout, hw, k, stride, res in LAYERS:
        if only and name not in only:
            continue

        class Net(nn.Module):
            def __init__(self):
                super().__init__()
                self.pre = nn.GroupConv2d(in_channels=8, out_channels=cin, kernel_size=1, padding=0, b_init=None)
                self.conv = nn.GroupConv2d(in_channels=cin, out_channels=cout, kernel_size=k, stride=stride, padding=(k - 1) // 2,
                                           b_init=None)
                self.bn = nn.BatchNorm2d(num_features=cout)
                self.act = nn.ReLU()

            def forward(self, x, r=None):
                y = self.bn(self.conv(self.pre(x)))
                if r is not None:
                    y = y + r
                return self.act(y)

        net = Net().cuda().set_eval()
        x = torch.randn(n, 8, hw, hw, device="cuda")
        po = (hw + 2 * ((k - 1) // 2) - k) // stride + 1
        args = (x, torch.randn(n, cout, po, po, device="cuda")) if res else (x,)
        results = []
        for combo in itertools.product(*ENV.values()):
            for key, val in zip(ENV, combo):
                if val is None:
                    os.environ.pop(key, None)
                else:
                    os.environ[key] = val
            net.invalidate_plans()
            try:
                plan, _, flat = runtime.get_plan(net, args, {})
            except Exception as e:  # a combination the kernel refuses
                continue
            outs = plan.alloc_outputs()
            for _ in range(2):
                plan.run(flat, outs, graph=False)
            torch.cuda.synchronize()
            best, kern = None, ""
            for _ in range(3):
                convs = [p for p in plan.profile(flat, outs) if p["path"] == "conv"]
                us = convs[-1]["ms"] * 1e3
                best = us if best is None else min(best, us)
                kern = convs[-1]["kernel"]
            results.append((combo, best, kern))
            del plan, outs
        for key in ENV:
            os.environ.pop(key, None)
        auto = next(r for r in results if r[0] == (None, None, None))
        line = f"{name:11s} auto {auto[1]:6.1f} us {auto[2]:30s}"
        seen = set()
        for combo, us, kern in sorted(results, key=lambda r: r[1]):
            if us < auto[1] * 0.98 and (kern, round(us, 0)) not in seen:
                seen.add((kern, round(us, 0)))
                line += f" | {us:5.1f} bn={combo[0]} 2sm={combo[1]} ring={combo[2]} {kern[13:]}"
        print(line[:400], flush=True)


if __name__ == "__main__":
    main()
