#!/usr/bin/env python
"""Time the ResNet stem conv alone (7x7 s2 3->64 @224, batch 256)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from tlxcv_b200 import nn, runtime

class Stem(nn.Module):
    def __init__(self):
        super().__init__()
        self.conv = nn.GroupConv2d(in_channels=3, out_channels=64, kernel_size=7, stride=2, padding=3, b_init=None)
        self.bn = nn.BatchNorm2d(num_features=64)
        self.relu = nn.ReLU()
    def forward(self, x):
        return self.relu(self.bn(self.conv(x)))

net = Stem().cuda().set_eval()
x = torch.randn(256, 3, 224, 224, device="cuda")
plan, _, flat = runtime.get_plan(net, (x,), {})
outs = plan.alloc_outputs()
for _ in range(3):
    prof = plan.profile(flat, outs)
print(os.environ.get("TLXCV_DEBUG_ABLATE", "0"), [(p["kernel"], round(p["ms"] * 1e3, 1)) for p in prof])
