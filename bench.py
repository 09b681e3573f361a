#!/usr/bin/env python
"""Headline benchmark: ResNet-50 bs256 224x224 bf16 inference, images/s (BASELINE.json).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference]

One process per GPU (torchrun env for N > 1).  Per-GPU batch is fixed at 256 (weak scaling: the
global batch is 256*N, sharded one contiguous block per rank, logits all-gathered over NCCL every
step).  A step = one forward of the sharded batch through tlxcv_b200 (+ the gather).

JSON line (rank 0): `value` = whole-job images/s with inputs resident in HBM; `e2e` = the same
through the host-buffer API (pinned host batch -> H2D -> forward -> gather -> D2H logits, double
buffered); `roofline` = the conv tcgen05 kernel family, FLOPs / CUDA-event time measured live;
`cpu_baseline` = the oracle restatement of the reference forward on this box's host cores.
`--impl reference` times that CPU path alone (tensorlayerx itself is not installable; the oracle
is the reference's model code over torch.nn.functional — see oracle/__init__.py).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

MODEL = "resnet50"
PER_GPU_BATCH = 256
SIZE = 224
METRIC = "resnet50_bs256_images_per_sec"
UNIT = "images/s"


TENSOR_KERNELS = ("conv_tcgen05", "conv3x3_slab", "stem_rowring")


def load_ncu_traffic():
    """DRAM bytes (read + write) of the tensor-core conv family per step, from the committed ncu pass
    (tools/ncu_round.sh -> tools/ncu_summary.py -> profiles/ncu_traffic.json); None when no capture is committed."""
    p = os.path.join(ROOT, "profiles", "ncu_traffic.json")
    try:
        return float(json.load(open(p))["conv_family_dram_bytes_per_step"])
    except Exception:  # noqa: BLE001
        return None


def load_peaks():
    peaks = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "source": "fallback"}
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        try:
            d = json.load(open(p))
            peaks.update(hbm_gbs=float(d["hbm_gbs"]), bf16_tflops=float(d["bf16_tflops"]), source="measured",
                         bf16_tflops_sustained=float(d.get("bf16_tflops_sustained", d["bf16_tflops"])))
        except Exception:  # noqa: BLE001
            pass
    return peaks


class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms during the timed region."""
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu, self.rows, self.proc = gpu_index, [], None

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-lms", "200", "-i", str(self.gpu)], stdout=subprocess.PIPE, text=True)
            threading.Thread(target=self._read, daemon=True).start()
        except Exception:  # noqa: BLE001
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.25)
        self.proc.terminate()
        sm, mx, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1])), mx.append(float(r[2]))
            except (ValueError, IndexError):
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), r[4:8]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


def cpu_forward_rate(n_images, min_seconds, max_iters=50):
    """Oracle restatement of the reference's ResNet-50 forward on the host cores (fp32, all threads)."""
    import torch

    from oracle import restated
    from tlxcv_b200.testing import seeded_state_dict, synthetic_images
    from tlxcv_b200 import models

    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    m = models.REGISTRY[MODEL]()
    sd = seeded_state_dict(m.state_dict(), MODEL)
    x = synthetic_images(n_images, SIZE)
    restated.forward(MODEL, sd, x[:2])                       # warm-up (oneDNN primitive creation)
    times = []
    t_end = time.time() + min_seconds
    while (time.time() < t_end or len(times) < 2) and len(times) < max_iters:
        t0 = time.perf_counter()
        restated.forward(MODEL, sd, x)
        times.append(time.perf_counter() - t0)
    best = min(times)
    return n_images / best, cores, len(times), sd, x


def gpu_library_rates(device, iters=10):
    """SURVEY 8(d) "GPU library baseline": the reference's model code (oracle restatement, torch.nn.functional) moved to
    the GPU as it is - what TL_BACKEND=torch does on a CUDA device - so cuDNN / cuBLAS run every layer unfused.
    (i) fp32 NCHW eager, the reference's literal configuration; (ii) bf16 channels_last with cuDNN autotune.
    A reported baseline like cpu_baseline: it never feeds the product path."""
    import torch

    from oracle import restated
    from tlxcv_b200 import models
    from tlxcv_b200.testing import seeded_state_dict, synthetic_images

    out = {}
    m = models.REGISTRY[MODEL]()
    sd32 = {k: v.to(device) for k, v in seeded_state_dict(m.state_dict(), MODEL).items()}
    x32 = synthetic_images(8, SIZE).repeat(PER_GPU_BATCH // 8, 1, 1, 1).to(device)
    torch.backends.cudnn.benchmark = True
    for name, sd, x in (("fp32_nchw_eager", sd32, x32),
                        ("bf16_channels_last_eager",
                         {k: (v.to(torch.bfloat16).contiguous(memory_format=torch.channels_last) if v.dim() == 4 else
                              v.to(torch.bfloat16)) for k, v in sd32.items()},
                         x32.to(torch.bfloat16).contiguous(memory_format=torch.channels_last))):
        for _ in range(3):
            restated.forward(MODEL, sd, x)
        torch.cuda.synchronize(device)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(iters):
            restated.forward(MODEL, sd, x)
        e1.record()
        torch.cuda.synchronize(device)
        ms = e0.elapsed_time(e1) / iters
        out[name] = {"value": PER_GPU_BATCH / ms * 1e3, "unit": UNIT, "ms_per_step": ms}
    out["what"] = ("oracle restatement of the reference ResNet-50 forward run by torch eager on this GPU (cuDNN/cuBLAS, one "
                   "library kernel per op, BatchNorm / ReLU / add unfused), bs256, device-resident input")
    return out


def run_reference(args, rank):
    """--impl reference: the reference's CPU path (oracle) on the host cores, same metric and config."""
    if rank != 0:
        return 0
    import torch

    from oracle import restated

    sample = 16
    rate, cores, _, sd, x = cpu_forward_rate(sample, 0.0, max_iters=2)
    for _ in range(max(0, args.warmup - 1)):
        restated.forward(MODEL, sd, x)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        restated.forward(MODEL, sd, x)
    dt = time.perf_counter() - t0
    value = sample * args.steps / dt
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dt / args.steps * 1e3, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"resnet50 224x224 forward, reference model code (oracle restatement) on torch CPU "
                               f"fp32; each step = a {sample}-image sample of the bs256 batch",
                   "per_gpu_batch": PER_GPU_BATCH, "torch_threads": torch.get_num_threads()},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port",
                         "sample": f"{sample} images per step x {args.steps} steps"},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


_RESULT_FD = None


def emit(line):
    """Write the one JSON line to the process's original stdout (see main())."""
    data = (json.dumps(line) + "\n").encode()
    if _RESULT_FD is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_RESULT_FD, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--gpu-library-baseline", action="store_true",
                    help="also time the reference's math on cuDNN/cuBLAS (torch eager) on this GPU: adds gpu_library_baseline")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)

    # stdout carries exactly ONE line, the JSON result: everything libraries print while the run lasts (NCCL's version
    # banner at NCCL_DEBUG >= VERSION, torch warnings) goes to stderr; emit() writes the result to the real stdout
    sys.stdout.flush()
    global _RESULT_FD
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    if args.impl == "reference":
        return run_reference(args, rank)

    import torch
    import torch.distributed as tdist

    from tlxcv_b200 import dist, models, runtime
    from tlxcv_b200.pipeline import HostPipeline
    from tlxcv_b200.testing import seeded_state_dict, synthetic_images

    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a B200: tlxcv_b200 has no CPU fallback (use --impl reference for the CPU path)")
    rank, local_rank, world = dist.init_from_env("cuda")
    if world != args.gpus:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}: launch with torchrun --nproc-per-node {args.gpus}")
    torch.cuda.set_device(local_rank)
    device = torch.device("cuda", local_rank)
    peaks = load_peaks()

    model = models.REGISTRY[MODEL]()
    model.load_state_dict(seeded_state_dict(model.state_dict(), MODEL))
    model = model.to(device).set_eval()
    base = synthetic_images(8, SIZE, seed=100 + rank)
    x_host = base.repeat(PER_GPU_BATCH // 8, 1, 1, 1).contiguous().pin_memory()
    x_dev = x_host.to(device)
    plan, _, _ = runtime.get_plan(model, (x_dev,), {})
    logits = plan.alloc_outputs()
    gathered = torch.empty((world * PER_GPU_BATCH, logits[0].shape[1]), dtype=torch.float32, device=device)

    def step():
        plan.run([x_dev], logits, graph=True)
        if world > 1:
            tdist.all_gather_into_tensor(gathered, logits[0])

    def fence():
        torch.cuda.synchronize(device)
        if world > 1:
            tdist.barrier()
        torch.cuda.synchronize(device)

    # ---- device-resident throughput --------------------------------------------------------
    for _ in range(args.warmup):
        step()
    fence()
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        step()
    e1.record()
    fence()
    ms = torch.tensor([e0.elapsed_time(e1) / args.steps], device=device)
    if world > 1:
        tdist.all_reduce(ms, op=tdist.ReduceOp.MAX)
    ms_per_step = float(ms.item())

    # ---- end to end through the host-buffer API --------------------------------------------
    gather = (lambda t: dist.gather_rows(t, world * PER_GPU_BATCH)) if world > 1 else None
    pipe = HostPipeline(model, tuple(x_host.shape), depth=2, device=device, gather=gather)
    out_host = [torch.empty((world * PER_GPU_BATCH, logits[0].shape[1]), dtype=torch.float32).pin_memory() for _ in range(2)]
    for i in range(args.warmup):
        pipe.submit(x_host, out_host[i % 2])
    pipe.synchronize()
    fence()
    t0 = torch.cuda.Event(enable_timing=True)
    t1 = torch.cuda.Event(enable_timing=True)
    t0.record(pipe.copy_stream)
    for i in range(args.steps):
        pipe.submit(x_host, out_host[i % 2])
    t1.record(pipe.compute_stream)
    pipe.synchronize()
    fence()
    e2e_ms = torch.tensor([t0.elapsed_time(t1) / args.steps], device=device)
    if world > 1:
        tdist.all_reduce(e2e_ms, op=tdist.ReduceOp.MAX)
    e2e_ok = bool(torch.equal(out_host[(args.steps - 1) % 2][rank * PER_GPU_BATCH:(rank + 1) * PER_GPU_BATCH],
                              logits[0].cpu()))
    # ---- the same end-to-end call with uint8 NHWC host batches: Normalize + ToTensor run inside
    #      the plan's input kernel, so a step moves 38.5 MB instead of 154 MB across PCIe; informational, `e2e` above is
    #      the reference-facing fp32 NCHW call
    from tlxcv_b200 import vision
    net8 = vision.Preprocessed(model, mean=(125.31, 122.95, 113.86), std=(62.99, 62.09, 66.70)).set_eval()
    u8_host = torch.randint(0, 256, (PER_GPU_BATCH, SIZE, SIZE, 3), dtype=torch.uint8,
                            generator=torch.Generator().manual_seed(200 + rank)).pin_memory()
    pipe8 = HostPipeline(net8, tuple(u8_host.shape), depth=2, device=device, gather=gather, dtype=torch.uint8)
    for i in range(args.warmup):
        pipe8.submit(u8_host, out_host[i % 2])
    pipe8.synchronize()
    fence()
    u0, u1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    u0.record(pipe8.copy_stream)
    for i in range(args.steps):
        pipe8.submit(u8_host, out_host[i % 2])
    u1.record(pipe8.compute_stream)
    pipe8.synchronize()
    fence()
    u8_ms = torch.tensor([u0.elapsed_time(u1) / args.steps], device=device)
    if world > 1:
        tdist.all_reduce(u8_ms, op=tdist.ReduceOp.MAX)
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant kernel family (conv on tcgen05), measured live --------------
    roof = None
    if rank == 0:
        prof = None
        for _ in range(3):
            prof = plan.profile([x_dev], logits)
        conv = [p for p in prof if p["kernel"].startswith(TENSOR_KERNELS)]
        flops = sum(p["flops"] for p in conv)
        conv_ms = sum(p["ms"] for p in conv)
        all_ms = sum(p["ms"] for p in prof)
        achieved = flops / (conv_ms * 1e-3) / 1e12
        roof = {"bound": "tensor", "achieved": achieved, "peak": peaks["bf16_tflops"], "unit": "TFLOP/s",
                "frac": achieved / peaks["bf16_tflops"], "traffic": load_ncu_traffic(), "peak_source": peaks["source"],
                "kernel": "tcgen05 conv family: conv_tcgen05_* / conv3x3_slab / stem_rowring (all conv + fc launches of a step)",
                "launches_per_step": len(conv),
                "flops_per_step": flops, "kernel_ms_per_step": conv_ms, "share_of_step": conv_ms / all_ms,
                "hbm_bound_layers": sum(1 for p in conv if p["bound"] == "hbm"),
                "algorithmic_bytes_per_step": sum(p["bytes"] for p in conv)}

    # ---- CPU baseline on this box's host cores (rank 0, single-GPU runs only) -----------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        sample = 16
        rate, cores, iters, _, _ = cpu_forward_rate(sample, 12.0)
        cpu = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
               "sample": f"best of {iters} forwards of a {sample}-image sample (oracle restatement of the reference "
                         f"model code, torch CPU fp32, {cores} threads)"}

    lib = None
    if rank == 0 and world == 1 and args.gpu_library_baseline:
        lib = gpu_library_rates(device)

    if rank == 0:
        total = world * PER_GPU_BATCH
        line = {
            "metric": METRIC, "value": total / ms_per_step * 1e3, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
            "config": {"workload": "resnet50 bs256/GPU 224x224 inference forward (random-init seeded weights)",
                       "per_gpu_batch": PER_GPU_BATCH, "global_batch": total, "parallelism": f"dp{world} (batch shard + "
                       "logits all-gather)" if world > 1 else "single GPU", "cuda_graph": True,
                       "l2": "inputs larger than L2: 154 MB input + ~11 GB activation traffic per step"},
            "e2e": {"value": total / float(e2e_ms.item()) * 1e3, "unit": UNIT, "h2d_bytes_per_step": pipe.h2d_bytes,
                    "d2h_bytes_per_step": out_host[0].numel() * 4, "ms_per_step": float(e2e_ms.item()),
                    "api": "tlxcv_b200.pipeline.HostPipeline (double-buffered)", "matches_device_path": e2e_ok},
            "e2e_uint8_input": {"value": total / float(u8_ms.item()) * 1e3, "unit": UNIT, "h2d_bytes_per_step": pipe8.h2d_bytes,
                                "d2h_bytes_per_step": out_host[0].numel() * 4, "ms_per_step": float(u8_ms.item()),
                                "api": "vision.Preprocessed(model) through HostPipeline: uint8 NHWC batches, normalisation fused "
                                       "into the plan's input kernel"},
            "gpu_launches": plan.num_launches * args.steps,
            "launches_per_step": plan.num_launches,
            "roofline": roof, "cpu_baseline": cpu, "clocks": clocks,
        }
        if lib is not None:
            line["gpu_library_baseline"] = lib
        emit(line)
    if world > 1:
        tdist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
