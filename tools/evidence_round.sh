#!/bin/bash
# Round-end evidence (run under gpurun, 1 GPU): GPU parity suite, bench + ncu passes (tools/ncu_round.sh), per-op profiles
# of every BASELINE config.  Everything lands in gpurun_out/; tools/ncu_summary.py + a copy into profiles/ follow on the host.
TAG=${1:-r02}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu_${TAG}.log 2>&1; echo "pytest rc=$?"
tail -2 gpurun_out/pytest_gpu_${TAG}.log
bash tools/ncu_round.sh ${TAG}; echo "ncu_round rc=$?"
cat gpurun_out/bench_${TAG}.json
prof() { timeout 300 python tools/quick_prof.py --model $1 --batch $2 --size $3 > gpurun_out/perop_${TAG}_$1_bs$2.txt 2>&1; head -1 gpurun_out/perop_${TAG}_$1_bs$2.txt; }
prof resnet50 256 224
prof resnext50_32x4d 256 224
prof mobilenet_v2 512 224
prof mobilenet_v1 512 224
prof darknet53_det 64 608
prof darknet53_cls 256 224
prof resnet18 256 224
prof yolov3_darknet53 64 608
prof mobilenet_v1_det 256 300
prof resnest50 256 224
