#!/usr/bin/env python
"""Key metrics of an `ncu --set full` report (tools/ncu_round.sh step 4), one row per launch.

    ncu -i gpurun_out/conv_full_TAG.ncu-rep --page raw --csv > /tmp/full.csv
    python tools/ncu_full_summary.py /tmp/full.csv > profiles/rNN_ncu_full_TAG.txt
"""
import csv
import sys

WANT = [
    ("gpu__time_duration.sum", "us"),
    ("launch__registers_per_thread", "regs"),
    ("launch__grid_size", "grid"),
    ("dram__bytes_read.sum", "dram_rd"),
    ("dram__bytes_write.sum", "dram_wr"),
    ("dram__throughput.avg.pct_of_peak_sustained_elapsed", "dram%"),
    ("lts__t_bytes.sum", "l2_bytes"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed", "tensor%"),
    ("sm__inst_executed_pipe_tmem.sum", "tmem_inst"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "sm%"),
    ("smsp__issue_active.avg.pct_of_peak_sustained_elapsed", "issue%"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smem_wavefronts"),
    ("smsp__cycles_active.avg", "cycles"),
]

rows = [r for r in csv.reader(open(sys.argv[1])) if r]
hdr, units = rows[0], rows[1]
name_i = hdr.index("Kernel Name")
cols = [(hdr.index(m), label) for m, label in WANT if m in hdr]
print("# ncu --set full, --clock-control none: first conv-family launches of a ResNet-50 bs256 forward (cold caches)")
print(f"{'kernel':58s} " + " ".join(f"{label:>14s}" for _, label in cols))
print(f"{'':58s} " + " ".join(f"{units[i][:14]:>14s}" for i, _ in cols))
for r in rows[2:]:
    k = r[name_i].replace("tlxcv::<unnamed>::", "").replace("(int)", "").replace("(bool)", "")
    print(f"{k[:58]:58s} " + " ".join(f"{r[i][:14]:>14s}" for i, _ in cols))
