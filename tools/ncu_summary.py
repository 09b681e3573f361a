#!/usr/bin/env python
"""Summarise ncu CSV logs (tools/ncu_round.sh) into a small text table for profiles/.

    python tools/ncu_summary.py gpurun_out/launches_TAG.csv [gpurun_out/traffic_TAG.csv] > profiles/rNN_ncu_TAG.txt
"""
from __future__ import annotations

import csv
import re
import sys
from collections import OrderedDict, defaultdict


def read(path):
    rows = []
    with open(path, newline="") as f:
        lines = [l for l in f if l.startswith('"')]
    for r in csv.DictReader(lines):
        rows.append(r)
    return rows


def short(name):
    m = re.search(r"(\w+)<", name) or re.search(r"(\w+)\(", name)
    base = m.group(1) if m else name[:40]
    t = re.search(r"<([^>]*)>", name)
    return base + ("<" + t.group(1).replace("(int)", "").replace(" ", "") + ">" if t else "")


def main():
    launches = read(sys.argv[1])
    per_id = OrderedDict()
    for r in launches:
        d = per_id.setdefault(r["ID"], {"kernel": short(r["Kernel Name"]), "grid": r["Grid Size"], "block": r["Block Size"]})
        d[r["Metric Name"]] = float(r["Metric Value"].replace(",", ""))
        d["unit_" + r["Metric Name"]] = r["Metric Unit"]
    print(f"# ncu launch list: {sys.argv[1]}  ({len(per_id)} launches; cold-cache serialised times: compare shares)")
    agg = defaultdict(lambda: [0, 0.0])
    tot = 0.0
    for d in per_id.values():
        t = d.get("gpu__time_duration.sum", 0.0)
        if d.get("unit_gpu__time_duration.sum", "ns") in ("us", "usecond"):
            t *= 1e3
        elif d.get("unit_gpu__time_duration.sum") in ("ms", "msecond"):
            t *= 1e6
        agg[d["kernel"]][0] += 1
        agg[d["kernel"]][1] += t
        tot += t
    print(f"{'kernel':70s} {'launches':>8s} {'total_us':>10s} {'share':>7s}")
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"{k[:70]:70s} {n:8d} {t / 1e3:10.1f} {t / tot * 100:6.1f}%")
    if len(sys.argv) > 2:
        tr = read(sys.argv[2])
        per = OrderedDict()
        for r in tr:
            d = per.setdefault(r["ID"], {"kernel": short(r["Kernel Name"]), "grid": r["Grid Size"]})
            v = float(r["Metric Value"].replace(",", ""))
            u = r["Metric Unit"]
            scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6}.get(u, 1.0)
            d[r["Metric Name"]] = v * scale
        print(f"\n# per-launch DRAM / L2 traffic: {sys.argv[2]} ({len(per)} launches)")
        print(f"{'id':>4s} {'kernel':52s} {'grid':>12s} {'us':>8s} {'dram_rd_MB':>10s} {'dram_wr_MB':>10s} {'l2_MB':>9s} {'tensor%':>8s}")
        for i, d in per.items():
            tp = [v for k, v in d.items() if k.startswith("sm__pipe_tensor")]
            print(f"{i:>4s} {d['kernel'][:52]:52s} {d['grid']:>12s} {d.get('gpu__time_duration.sum', 0) / 1e3:8.1f} "
                  f"{d.get('dram__bytes_read.sum', 0) / 1e6:10.1f} {d.get('dram__bytes_write.sum', 0) / 1e6:10.1f} "
                  f"{d.get('lts__t_bytes.sum', 0) / 1e6:9.1f} {(tp[0] if tp else 0):8.1f}")


def traffic_json(path, out_json):
    """Per-step DRAM traffic of the tensor-core conv family from the traffic pass (-k filter; 4 forwards of one plan)."""
    import json
    tr = read(path)
    per = OrderedDict()
    for r in tr:
        d = per.setdefault(r["ID"], {"kernel": short(r["Kernel Name"])})
        v = float(r["Metric Value"].replace(",", ""))
        scale = {"Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(r["Metric Unit"], 1.0)
        d[r["Metric Name"]] = v * scale
    fam = ("conv_tcgen05", "conv_chain", "conv3x3_slab", "stem_rowring")
    ids = list(per.values())
    # one step = the launches between two consecutive import kernels
    starts = [i for i, d in enumerate(ids) if d["kernel"].startswith("import_nchw")]
    if len(starts) < 2:
        return
    step = ids[starts[0]:starts[1]]
    conv = [d for d in step if d["kernel"].startswith(fam)]
    tot = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in conv)
    allb = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in step)
    json.dump({"source": path, "launches_per_step": len(step), "conv_family_launches_per_step": len(conv),
               "conv_family_dram_bytes_per_step": tot, "all_kernels_dram_bytes_per_step": allb,
               "note": "dram__bytes_read.sum + dram__bytes_write.sum per launch (ncu, cold caches, serialised), summed over "
                       "the conv-family launches of one forward"}, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main()
    if len(sys.argv) > 3:
        traffic_json(sys.argv[2], sys.argv[3])
