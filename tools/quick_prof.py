#!/usr/bin/env python
"""Per-op device-time profile of one model plan (CUDA events between ops), plus whole-forward timing.

    python tools/quick_prof.py --model resnet50 --batch 256 --size 224 [--out gpurun_out/prof_resnet50.json]
"""
from __future__ import annotations

import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    import torch

    from tlxcv_b200 import models, runtime
    from tlxcv_b200.testing import seeded_state_dict, synthetic_images

    ap = argparse.ArgumentParser()
    ap.add_argument("--model", default="resnet50")
    ap.add_argument("--batch", type=int, default=256)
    ap.add_argument("--size", type=int, default=224)
    ap.add_argument("--iters", type=int, default=20)
    ap.add_argument("--out", default="")
    a = ap.parse_args()
    peaks = {"hbm_gbs": 6546.6, "bf16_tflops": 1624.3}
    pk = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(pk):
        peaks.update(json.load(open(pk)))

    model = models.REGISTRY[a.model]()
    model.load_state_dict(seeded_state_dict(model.state_dict(), a.model))
    model = model.cuda().set_eval()
    x = synthetic_images(min(a.batch, 8), a.size).cuda()
    x = x.repeat((a.batch + x.shape[0] - 1) // x.shape[0], 1, 1, 1)[:a.batch].contiguous()
    from tlxcv_b200.testing import DICT_INPUT
    args = ({"images": x},) if a.model in DICT_INPUT else (x,)
    plan, _, flat = runtime.get_plan(model, args, {})
    outs = plan.alloc_outputs()
    for _ in range(3):
        plan.run(flat, outs, graph=True)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(a.iters):
        plan.run(flat, outs, graph=True)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / a.iters
    prof = None
    for _ in range(3):
        prof = plan.profile(flat, outs)
    tot = sum(p["ms"] for p in prof)
    print(f"{a.model} bs{a.batch} {a.size}^2: {ms:.3f} ms/forward (graph) = {a.batch / ms * 1e3:.0f} img/s; "
          f"sum of per-op {tot:.3f} ms; launches {plan.num_launches}; workspace {plan.workspace_bytes / 2**20:.0f} MiB")
    print(f"{'idx':>3} {'path':38s} {'kernel':34s} {'ms':>8s} {'TF/s':>8s} {'GB/s':>8s} {'frac':>6s} bound grid")
    for p in prof:
        t = p["ms"] * 1e-3
        tf = p["flops"] / t / 1e12 if t > 0 else 0
        gb = p["bytes"] / t / 1e9 if t > 0 else 0
        frac = max(tf / peaks["bf16_tflops"], gb / peaks["hbm_gbs"])
        p.update(tflops=tf, gbs=gb, roofline_frac=frac)
        print(f"{p['index']:3d} {p['path'][:38]:38s} {p['kernel'][:34]:34s} {p['ms']:8.4f} {tf:8.1f} {gb:8.0f} {frac:6.2f} "
              f"{p['bound'][:4]:5s} {p['grid']}")
    conv = [p for p in prof if p["op"] in ("conv", "linear")]
    cflops, cms = sum(p["flops"] for p in conv), sum(p["ms"] for p in conv)
    print(f"conv+fc: {cflops / 1e12:.3f} TFLOP in {cms:.3f} ms = {cflops / cms / 1e9:.1f} TFLOP/s "
          f"({cflops / cms / 1e9 / peaks['bf16_tflops'] * 100:.1f}% of measured bf16 peak)")
    if a.out:
        os.makedirs(os.path.dirname(os.path.abspath(a.out)), exist_ok=True)
        json.dump(dict(model=a.model, batch=a.batch, size=a.size, ms_per_forward=ms, img_per_s=a.batch / ms * 1e3,
                       launches=plan.num_launches, ops=prof), open(a.out, "w"), indent=1)


if __name__ == "__main__":
    main()
