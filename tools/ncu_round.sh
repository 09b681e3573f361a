#!/bin/bash
# Evidence pass for one round (run under gpurun, 1 GPU):
#   1. plain bench run (must exit 0)                     -> gpurun_out/bench_TAG.json
#   2. ncu launch list of the same short command           -> gpurun_out/launches_TAG.csv
#   3. ncu DRAM/L2 traffic + tensor-pipe activity per launch -> gpurun_out/traffic_TAG.csv
#   4. ncu --set full of the first 8 conv launches         -> gpurun_out/conv_full_TAG.ncu-rep
set -u
TAG=${1:-r01}
CMD="python bench.py --steps 2 --warmup 3 --quick"
mkdir -p gpurun_out
python bench.py --steps 100 --warmup 10 > gpurun_out/bench_${TAG}.json 2> gpurun_out/bench_${TAG}.err || exit 1
$CMD > gpurun_out/plain_${TAG}.log 2>&1 || exit 2
ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/launches_${TAG}.csv $CMD > gpurun_out/ncu1_${TAG}.log 2>&1
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed --clock-control none -k regex:'conv_tcgen05|conv_chain|dwconv|maxpool|gap_|import_|export_|argmax|add_act|stem|slab' -c 260 --csv --log-file gpurun_out/traffic_${TAG}.csv $CMD > gpurun_out/ncu2_${TAG}.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"conv_tcgen05|conv_chain|stem_rowring|conv3x3_slab" -c 12 -o gpurun_out/conv_full_${TAG} -f $CMD > gpurun_out/ncu3_${TAG}.log 2>&1
echo done
