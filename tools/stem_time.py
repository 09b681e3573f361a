#!/usr/bin/env python
"""Time the stem (+pool) of one model family alone: import + stem [+ maxpool].   python tools/stem_time.py resnet|mobilenet|darknet"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tlxcv_b200 import nn, runtime

kind = sys.argv[1] if len(sys.argv) > 1 else "resnet"
cfg = {"resnet": (256, 224, 64, 7, 2, 3, True), "mobilenet": (512, 224, 32, 3, 2, 1, False), "darknet": (64, 608, 32, 3, 1, 1, False)}[kind]
n, hw, cout, k, stride, pad, pool = cfg


class Net(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.GroupConv2d(in_channels=3, out_channels=cout, kernel_size=k, stride=stride, padding=pad, b_init=None)
        self.bn = nn.BatchNorm2d(num_features=cout)
        self.act = nn.ReLU()
        self.pool = nn.MaxPool2d(kernel_size=3, stride=2, padding=1) if pool else None
        self.next = nn.GroupConv2d(in_channels=cout, out_channels=8, kernel_size=1, stride=1, padding=0, b_init=None)

    def forward(self, x):
        y = self.act(self.bn(self.conv(x)))
        y = self.pool(y) if self.pool is not None else y
        return self.next(y)


net = Net().cuda().set_eval()
x = torch.randn(n, 3, hw, hw, device="cuda")
plan, _, flat = runtime.get_plan(net, (x,), {})
outs = plan.alloc_outputs()
for _ in range(3):
    plan.run(flat, outs, graph=False)
torch.cuda.synchronize()
best = {}
for _ in range(5):
    for p in plan.profile(flat, outs):
        best[p["kernel"]] = min(best.get(p["kernel"], 1e9), p["ms"])
print(kind, os.environ.get("TLXCV_DEBUG_ABLATE_STEM", "0"), {k: round(v * 1e3, 1) for k, v in best.items()})
