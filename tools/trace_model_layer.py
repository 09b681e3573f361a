#!/usr/bin/env python
"""Dump the pipeline timeline of ONE tcgen05 conv launch of a whole-model forward.
    TLXCV_DEBUG_TRACE_CONV=out.bin TLXCV_DEBUG_TRACE_CONV_INDEX=k python tools/trace_model_layer.py resnet50 256
(k counts conv_tcgen05 launches from process start; the first eager forward is launches 0..n-1)"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from tlxcv_b200 import models, runtime
from tlxcv_b200.testing import seeded_state_dict, synthetic_images

name, batch = sys.argv[1], int(sys.argv[2])
model = models.REGISTRY[name]()
model.load_state_dict(seeded_state_dict(model.state_dict(), name))
model = model.cuda().set_eval()
x = synthetic_images(8, 224).cuda().repeat(batch // 8, 1, 1, 1).contiguous()
plan, _, flat = runtime.get_plan(model, (x,), {})
outs = plan.alloc_outputs()
plan.run(flat, outs, graph=False)
torch.cuda.synchronize()
k = 0
for i in range(len(plan.spec.ops)):
    info = plan.op_info(i)
    if info["kernel"].startswith("conv_tcgen05"):
        print(k, plan.spec.ops[i].path, info["kernel"])
        k += 1
