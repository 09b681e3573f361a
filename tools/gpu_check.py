#!/usr/bin/env python
"""GPU bring-up checker: runs kernel / model parity cases one by one in a worker process so that a
trapping or hanging kernel costs one case, not the run.  Writes gpurun_out/check.log.

    python tools/gpu_check.py [--filter substr] [--timeout 180]
"""
from __future__ import annotations

import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


# ------------------------------------------------------------------------------------------------
# case table
# ------------------------------------------------------------------------------------------------
def conv_case(name, n, cin, hw, cout, k, stride, pad, groups=1, bn=True, act=None, res=False, act2=None, prec="bf16",
              bias=False):
    return dict(kind="conv", name=name, n=n, cin=cin, hw=hw, cout=cout, k=k, stride=stride, pad=pad, groups=groups,
                bn=bn, act=act, res=res, act2=act2, prec=prec, bias=bias)


CASES = [
    # plumbing + every memory-bound kernel through the fp32 validation mode first
    conv_case("f32_1x1", 2, 64, 14, 64, 1, 1, 0, prec="f32", act="relu"),
    conv_case("f32_3x3_res", 2, 32, 9, 32, 3, 1, 1, prec="f32", act="leaky", res=True),
    conv_case("f32_stem7", 2, 3, 32, 64, 7, 2, 3, prec="f32", act="relu"),
    conv_case("f32_dw", 2, 32, 12, 32, 3, 2, 1, groups=32, prec="f32", act="relu6"),
    conv_case("f32_grouped", 2, 128, 8, 128, 3, 1, 1, groups=32, prec="f32", act="relu"),
    # tcgen05: plain GEMM tiles
    conv_case("tc_1x1_c64_k64", 2, 64, 16, 64, 1, 1, 0, act=None, bn=False),
    conv_case("tc_1x1_c64_k64_bn_relu", 2, 64, 16, 64, 1, 1, 0, act="relu"),
    conv_case("tc_1x1_c256_k128", 3, 256, 14, 128, 1, 1, 0, act="relu"),
    conv_case("tc_1x1_c128_k512_res", 2, 128, 14, 512, 1, 1, 0, res=True, act2="relu"),
    conv_case("tc_1x1_mtail", 1, 64, 7, 256, 1, 1, 0, act="relu"),
    conv_case("tc_1x1_c24_k144", 2, 24, 12, 144, 1, 1, 0, act="relu6"),
    conv_case("tc_1x1_c96_k24", 2, 96, 12, 24, 1, 1, 0),
    # tcgen05: im2col TMA
    conv_case("tc_3x3_c64", 2, 64, 14, 64, 3, 1, 1, act="relu"),
    conv_case("tc_3x3_c128_s2", 2, 128, 28, 128, 3, 2, 1, act="relu"),
    conv_case("tc_3x3_c256_7x7", 5, 256, 7, 256, 3, 1, 1, act="relu"),
    conv_case("tc_1x1_s2_ds", 2, 256, 14, 512, 1, 2, 0),
    conv_case("tc_3x3_c32_s2_leaky", 2, 32, 20, 64, 3, 2, 1, act="leaky"),
    conv_case("tc_3x3_leaky_res", 2, 64, 10, 128, 3, 1, 1, act="leaky", res=True),
    # grouped (block-diagonal) and depthwise
    conv_case("tc_grouped_c128_g32", 2, 128, 14, 128, 3, 1, 1, groups=32, act="relu"),
    conv_case("tc_grouped_c256_g32_s2", 2, 256, 14, 256, 3, 2, 1, groups=32, act="relu"),
    conv_case("dw_c32_s1", 2, 32, 14, 32, 3, 1, 1, groups=32, act="relu6"),
    conv_case("dw_c96_s2", 2, 96, 15, 96, 3, 2, 1, groups=96, act="relu6"),
    # stems (C_in = 3 gather producer)
    conv_case("stem_7x7_s2", 2, 3, 64, 64, 7, 2, 3, act="relu"),
    conv_case("stem_3x3_s2", 2, 3, 32, 32, 3, 2, 1, act="relu6"),
    conv_case("stem_3x3_s1", 2, 3, 24, 32, 3, 1, 1, act="leaky"),
    # models
    dict(kind="model", name="m_resnet50_f32", model="resnet50", n=2, size=64, prec="f32"),
    dict(kind="model", name="m_resnet50_bf16", model="resnet50", n=4, size=224, prec="bf16"),
    dict(kind="model", name="m_resnet18_bf16", model="resnet18", n=2, size=224, prec="bf16"),
    dict(kind="model", name="m_resnext50_bf16", model="resnext50_32x4d", n=2, size=224, prec="bf16"),
    dict(kind="model", name="m_mobilenet_v2_bf16", model="mobilenet_v2", n=2, size=224, prec="bf16"),
    dict(kind="model", name="m_mobilenet_v1_bf16", model="mobilenet_v1", n=2, size=224, prec="bf16"),
    dict(kind="model", name="m_darknet53_cls_bf16", model="darknet53_cls", n=2, size=224, prec="bf16"),
    dict(kind="model", name="m_darknet53_det_bf16", model="darknet53_det", n=1, size=64, prec="bf16"),
    dict(kind="model", name="m_resnet50_bf16_golden", model="resnet50", n=4, size=224, prec="bf16", golden=True),
]


# ------------------------------------------------------------------------------------------------
# worker
# ------------------------------------------------------------------------------------------------
def run_conv_case(c):
    import torch
    import torch.nn.functional as F

    import tlxcv_b200 as tlx
    from tlxcv_b200 import nn, runtime

    g = torch.Generator().manual_seed(hash(c["name"]) % 10007)
    cin, cout, k, groups = c["cin"], c["cout"], c["k"], c["groups"]
    x = torch.randn(c["n"], cin, c["hw"], c["hw"], generator=g)
    w = torch.randn(cout, cin // groups, k, k, generator=g) * (2.0 / (cin // groups * k * k)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1 if c["bias"] else None
    gamma = 0.75 + 0.5 * torch.rand(cout, generator=g)
    beta = torch.randn(cout, generator=g) * 0.1
    mean = torch.randn(cout, generator=g) * 0.1
    var = 0.75 + 0.5 * torch.rand(cout, generator=g)
    p_out = (c["hw"] + 2 * c["pad"] - k) // c["stride"] + 1
    res = torch.randn(c["n"], cout, p_out, p_out, generator=g) if c["res"] else None
    acts = {None: None, "relu": nn.ReLU, "relu6": nn.ReLU6, "leaky": lambda: nn.LeakyReLU(0.1)}

    class Net(nn.Module):
        def __init__(self):
            super().__init__()
            self.conv = nn.GroupConv2d(in_channels=cin, out_channels=cout, kernel_size=k, stride=c["stride"],
                                       padding=c["pad"], n_group=groups, b_init="constant" if c["bias"] else None)
            self.bn = nn.BatchNorm2d(num_features=cout) if c["bn"] else None
            self.a1 = acts[c["act"]]() if c["act"] else None
            self.a2 = acts[c["act2"]]() if c["act2"] else None

        def forward(self, x, r=None):
            y = self.conv(x)
            if self.bn is not None:
                y = self.bn(y)
            if self.a1 is not None:
                y = self.a1(y)
            if r is not None:
                y = y + r
            if self.a2 is not None:
                y = self.a2(y)
            return y

    net = Net()
    sd = {"conv.filters": w}
    if b is not None:
        sd["conv.biases"] = b
    if c["bn"]:
        sd.update({"bn.beta": beta, "bn.gamma": gamma, "bn.moving_mean": mean, "bn.moving_var": var})
    net.load_state_dict(sd)
    net = net.cuda().set_eval()
    prec = runtime.PREC_F32 if c["prec"] == "f32" else runtime.PREC_BF16
    args = (x.cuda(),) if res is None else (x.cuda(), res.cuda())
    plan, structure, flat = runtime.get_plan(net, args, {}, precision=prec)
    out = plan.run([t.contiguous() for t in flat], graph=False)[0].cpu()
    torch.cuda.synchronize()

    def q(t):
        return t if c["prec"] == "f32" else t.bfloat16().float()

    # reference: fp32 math on (bf16-rounded in bf16 mode) inputs and weights
    y = F.conv2d(q(x), q(w), b, c["stride"], c["pad"], 1, groups)
    if c["bn"]:
        y = F.batch_norm(y, mean, var, gamma, beta, False, 0.0, 1e-5)
    fa = {None: lambda t: t, "relu": F.relu, "relu6": F.relu6, "leaky": lambda t: F.leaky_relu(t, 0.1)}
    y = fa[c["act"]](y)
    if res is not None:
        y = y + q(res)
    y = fa[c["act2"]](y)
    err = (out - y).abs()
    scale = float(y.abs().max())
    tol = 2e-4 * max(scale, 1.0) if c["prec"] == "f32" else (2.0 ** -7) * max(scale, 1.0)
    info = dict(max_err=float(err.max()), ref_max=scale, tol=tol, kernels=[plan.op_info(i)["kernel"] for i in range(len(plan.spec.ops))])
    if not bool(torch.isfinite(out).all()):
        info["nonfinite"] = int((~torch.isfinite(out)).sum())
    ok = float(err.max()) <= tol and "nonfinite" not in info
    if not ok:
        # localise: which output positions / channels are wrong
        bad = (err > tol)
        info["bad_frac"] = float(bad.float().mean())
        idx = bad.nonzero()[:6].tolist()
        info["bad_idx"] = idx
        info["bad_vals"] = [(float(out[tuple(i)]), float(y[tuple(i)])) for i in idx]
        info["bad_per_channel_frac"] = [round(float(v), 3) for v in bad.float().mean(dim=(0, 2, 3))[:16]]
        info["bad_rows_first_img"] = [round(float(v), 2) for v in bad[0].float().mean(dim=(0, 2))[:16]]
    return ok, info


def run_model_case(c):
    import numpy as np
    import torch

    from oracle import restated
    from tlxcv_b200 import models, runtime
    from tlxcv_b200.testing import seeded_state_dict, synthetic_images

    name = c["model"]
    model = models.REGISTRY[name]()
    sd = seeded_state_dict(model.state_dict(), name)
    model.load_state_dict(sd)
    model = model.cuda().set_eval()
    x = synthetic_images(c["n"], c["size"])
    is_det = name == "darknet53_det"
    ref = restated.forward(name, sd, {"images": x} if is_det else x)
    refs = ref if isinstance(ref, list) else [ref]
    if c.get("golden"):
        g = np.load(os.path.join(ROOT, "tests/golden", f"{name}.npz"))
        refs = [torch.from_numpy(g[f"out{i}"]) for i in range(len(refs))]
    prec = runtime.PREC_F32 if c["prec"] == "f32" else runtime.PREC_BF16
    xin = x.cuda()
    args = ({"images": xin},) if is_det else (xin,)
    plan, structure, flat = runtime.get_plan(model, args, {}, precision=prec)
    outs = [o.cpu() for o in plan.run(flat, graph=False)]
    outs2 = [o.cpu() for o in plan.run(flat, graph=True)]
    outs3 = [o.cpu() for o in plan.run(flat, graph=True)]
    torch.cuda.synchronize()
    info = dict(launches=plan.num_launches, workspace_mb=plan.workspace_bytes / 2 ** 20)
    ok = True
    for i, (o, r) in enumerate(zip(outs, refs)):
        err = float((o - r).abs().max())
        info[f"out{i}_max_err"] = err
        info[f"out{i}_ref_std"] = float(r.std())
        info[f"out{i}_graph_equal"] = bool(torch.equal(o, outs2[i]) and torch.equal(o, outs3[i]))
        ok = ok and info[f"out{i}_graph_equal"] and bool(torch.isfinite(o).all())
        if not is_det:
            top2 = r.topk(2, dim=1).values
            margin = top2[:, 0] - top2[:, 1]
            agree = o.argmax(1) == r.argmax(1)
            info["argmax_agree"] = f"{int(agree.sum())}/{len(agree)}"
            info["min_margin"] = float(margin.min())
            tol = 1e-4 if c["prec"] == "f32" else 1e-2
            ok = ok and err <= tol and bool((agree | (margin < 2 * tol)).all())
        else:
            tol = (1e-4 if c["prec"] == "f32" else 0.03) * max(1.0, float(r.abs().max()))
            ok = ok and err <= tol
        info[f"out{i}_tol"] = tol
    return ok, info


def worker(names):
    import traceback

    cases = {c["name"]: c for c in CASES}
    for nm in names:
        c = cases[nm]
        print(f"BEGIN {nm}", flush=True)
        t0 = time.time()
        try:
            ok, info = run_conv_case(c) if c["kind"] == "conv" else run_model_case(c)
            status = "PASS" if ok else "FAIL"
        except Exception as e:  # noqa: BLE001
            status, info = "ERROR", dict(error=f"{type(e).__name__}: {e}"[:300], tb=traceback.format_exc()[-400:])
        info["secs"] = round(time.time() - t0, 2)
        print(f"END {nm} {status} {json.dumps(info)}", flush=True)
        if status == "ERROR" and "CUDA" in info.get("error", ""):
            sys.exit(3)      # context is likely dead: let the parent restart us


# ------------------------------------------------------------------------------------------------
# parent
# ------------------------------------------------------------------------------------------------
def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--worker", nargs="*")
    ap.add_argument("--filter", default="")
    ap.add_argument("--timeout", type=float, default=240.0)
    ap.add_argument("--log", default=os.path.join(ROOT, "gpurun_out", "check.log"))
    a = ap.parse_args()
    if a.worker is not None:
        return worker(a.worker)
    os.makedirs(os.path.dirname(a.log), exist_ok=True)
    pending = [c["name"] for c in CASES if a.filter in c["name"]]
    results = {}
    log = open(a.log, "a")

    def emit(s):
        print(s, flush=True)
        log.write(s + "\n")
        log.flush()

    emit(f"=== gpu_check {time.strftime('%F %T')} cases={len(pending)}")
    while pending:
        proc = subprocess.Popen([sys.executable, os.path.abspath(__file__), "--worker", *pending], stdout=subprocess.PIPE,
                                stderr=subprocess.STDOUT, text=True, cwd=ROOT)
        current = [None]
        last = [time.time()]

        def watchdog():
            while proc.poll() is None:
                if time.time() - last[0] > a.timeout:
                    proc.kill()
                    return
                time.sleep(1.0)

        threading.Thread(target=watchdog, daemon=True).start()
        tail = []
        for line in proc.stdout:
            line = line.rstrip("\n")
            last[0] = time.time()
            if line.startswith("BEGIN "):
                current[0] = line.split()[1]
                tail = []
            elif line.startswith("END "):
                _, nm, status, info = line.split(" ", 3)
                results[nm] = status
                pending.remove(nm)
                current[0] = None
                emit(f"{status:5s} {nm} {info}")
            else:
                tail.append(line)
                tail = tail[-15:]
        proc.wait()
        if current[0] is not None:
            nm = current[0]
            results[nm] = "CRASH"
            pending.remove(nm)
            emit(f"CRASH {nm} rc={proc.returncode} tail={json.dumps(tail[-8:])}")
        elif proc.returncode not in (0, 3) and pending:
            emit(f"worker exited rc={proc.returncode} tail={json.dumps(tail[-8:])}")
            nm = pending.pop(0)
            results[nm] = "CRASH"
    n_pass = sum(1 for v in results.values() if v == "PASS")
    emit(f"=== {n_pass}/{len(results)} passed; failures: {[k for k, v in results.items() if v != 'PASS']}")
    return 0 if n_pass == len(results) else 1


if __name__ == "__main__":
    sys.exit(main())
