#!/usr/bin/env python
"""Timing experiments on single conv layers: which resource bounds the tcgen05 conv kernel?

TLXCV_DEBUG_ABLATE bits: 1 gather loads, 2 output stores, 4 MMAs, 8 operand loads, 16 residual loads,
32 epilogue math + store.  Results are wrong with any bit set; timing only.  One process, one plan per
(layer, ablation) pair; prints microseconds per variant.
"""
from __future__ import annotations

import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

LAYERS = [  # name, cin, cout, hw, k, stride, res, env
    ("l1conv3", 64, 256, 56, 1, 1, True, {}),
    ("l2conv3", 128, 512, 28, 1, 1, True, {}),
    ("l3conv3", 256, 1024, 14, 1, 1, True, {}),
    ("l3conv1", 1024, 256, 14, 1, 1, False, {}),
    ("l3conv1_nopair", 1024, 256, 14, 1, 1, False, {"TLXCV_DEBUG_2SM": "0"}),
    ("l4conv3", 512, 2048, 7, 1, 1, True, {}),
    ("l4conv1", 2048, 512, 7, 1, 1, False, {}),
    ("l3conv2", 256, 256, 14, 3, 1, False, {}),
    ("l3conv2_nopair", 256, 256, 14, 3, 1, False, {"TLXCV_DEBUG_2SM": "0"}),
    ("l2conv2", 128, 128, 28, 3, 1, False, {}),
    ("l2conv2_pair", 128, 128, 28, 3, 1, False, {"TLXCV_DEBUG_2SM": "1"}),
    ("l2conv1", 512, 128, 28, 1, 1, False, {}),
    ("l2conv1_pair", 512, 128, 28, 1, 1, False, {"TLXCV_DEBUG_2SM": "1"}),
    ("l4conv2", 512, 512, 7, 3, 1, False, {}),
    ("l4conv2_pair", 512, 512, 7, 3, 1, False, {"TLXCV_DEBUG_2SM": "1"}),
    ("l2conv2_s2", 128, 128, 28, 3, 1, False, {"TLXCV_DEBUG_STAGES": "2"}),
    ("l2conv2_s3", 128, 128, 28, 3, 1, False, {"TLXCV_DEBUG_STAGES": "3"}),
    ("l2conv2_s4", 128, 128, 28, 3, 1, False, {"TLXCV_DEBUG_STAGES": "4"}),
    ("l2conv2_r2", 128, 128, 28, 3, 1, False, {"TLXCV_DEBUG_RING": "2"}),
    ("l2conv2_pair_r2", 128, 128, 28, 3, 1, False, {"TLXCV_DEBUG_2SM": "1", "TLXCV_DEBUG_RING": "2"}),
    ("l3conv2_s2", 256, 256, 14, 3, 1, False, {"TLXCV_DEBUG_STAGES": "2"}),
    ("l3conv2_s3", 256, 256, 14, 3, 1, False, {"TLXCV_DEBUG_STAGES": "3"}),
    ("l3conv2_r4", 256, 256, 14, 3, 1, False, {"TLXCV_DEBUG_RING": "4"}),
    # DarkNet-53 at 608x608, batch 64 ("_n")
    ("dn_ds0", 32, 64, 608, 3, 2, False, {"_n": "64"}),
    ("dn_ds0_kb64", 32, 64, 608, 3, 2, False, {"_n": "64", "TLXCV_NO_KB32": "1"}),
    ("dn_ds1", 64, 128, 304, 3, 2, False, {"_n": "64"}),
    ("dn_s1_3x3", 64, 128, 152, 3, 1, False, {"_n": "64"}),
]
ABLATIONS = [int(x) for x in os.environ.get("ABLATIONS", "0,2,16,18,8,24,26,32,40,56,4,12").split(",")]


def main():
    import torch

    import tlxcv_b200 as tlx  # noqa: F401
    from tlxcv_b200 import nn, runtime

    n = 256
    only = sys.argv[1:]
    for name, cin, cout, hw, k, stride, res, env in LAYERS:
        if only and name not in only:
            continue
        row = []
        for ab in ABLATIONS:
            if (ab & 16) and not res:
                continue
            os.environ["TLXCV_DEBUG_ABLATE"] = str(ab)
            print(f"[{name} ablate={ab}]", file=sys.stderr, flush=True)
            n = int(env.get("_n", 256))
            for kk, vv in env.items():
                os.environ[kk] = vv

            class Net(nn.Module):
                def __init__(self):
                    super().__init__()
                    self.pre = nn.GroupConv2d(in_channels=8, out_channels=cin, kernel_size=1, padding=0, b_init=None)
                    self.conv = nn.GroupConv2d(in_channels=cin, out_channels=cout, kernel_size=k, stride=stride,
                                               padding=(k - 1) // 2, b_init=None)
                    self.bn = nn.BatchNorm2d(num_features=cout)
                    self.act = nn.ReLU()

                def forward(self, x, r=None):
                    y = self.bn(self.conv(self.pre(x)))
                    if r is not None:
                        y = y + r
                    return self.act(y)

            net = Net().cuda().set_eval()
            x = torch.randn(n, 8, hw, hw, device="cuda")
            args = (x, torch.randn(n, cout, hw // stride, hw // stride, device="cuda")) if res else (x,)
            plan, _, flat = runtime.get_plan(net, args, {})
            outs = plan.alloc_outputs()
            for _ in range(3):
                plan.run(flat, outs, graph=False)
            torch.cuda.synchronize()
            best = None
            for _ in range(3):
                prof = plan.profile(flat, outs)
                convs = [p for p in prof if p["path"] == "conv"]
                us = convs[-1]["ms"] * 1e3
                best = us if best is None else min(best, us)
            row.append(f"{ab}:{best:.1f}")
            kern = convs[-1]["kernel"]
            stages = (convs[-1]["smem"] - 70000) // (49152 if "n256" in kern else 32768 if "n128" in kern else 24576) if "2sm" not in kern else -1
            del plan, outs, net
            for kk in env:
                os.environ.pop(kk, None)
        print(f"{name:16s} {kern:28s} st{stages} " + "  ".join(row), flush=True)
    os.environ.pop("TLXCV_DEBUG_ABLATE", None)


if __name__ == "__main__":
    main()
