#!/usr/bin/env python
"""Per-round timeline of conv_tcgen05 CTA 0 (fine-trace debug build, TLXCV_DEBUG_ABLATE bit 512, TLXCV_DEBUG_TRACE_CONV=<file>).

producer events per K block: before the empty-slot wait, slot free; MMA warp per round: before the full wait, operands
landed, issued + committed, (K blocks in the round)."""
import sys
import numpy as np

L = 4096
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(3, L).astype(np.int64)
prod = a[0][a[0] > 0]
mm = a[1][:(np.count_nonzero(a[1]) // 4) * 4].reshape(-1, 4)
first = int(np.argmax(a[1][:len(mm) * 4].reshape(-1, 4)[:, 0] > 1000))
t0 = prod[0]
prod = (prod[:len(prod) // 2 * 2] - t0).reshape(-1, 2)
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (40, 70)
print("producer K block: wait_start slot_free (wait)   | MMA round: wait_start landed issued kbs (wait, issue)")
kbi = 0
for r in range(len(mm)):
    if r < first:
        kbi += int(mm[r, 3])
        continue
    if lo <= r < hi:
        w0, w1, w2, n = mm[r]
        ps = " ; ".join(f"{prod[kbi + j, 0]:7d} {prod[kbi + j, 1]:7d} ({prod[kbi + j, 1] - prod[kbi + j, 0]:4d})" for j in range(int(n)) if kbi + j < len(prod))
        print(f"{r:4d} | {w0 - t0:7d} {w1 - t0:7d} {w2 - t0:7d} {n} ({w1 - w0:4d}, {w2 - w1:4d}) | producer kb {kbi}: {ps}")
    kbi += int(mm[r, 3])
s = slice(max(20, first + 2), len(mm) - 5)
print(f"MMA rounds: mean period {np.mean(np.diff(mm[s, 0])):.0f}, wait {np.mean(mm[s, 1] - mm[s, 0]):.0f}, issue+commit {np.mean(mm[s, 2] - mm[s, 1]):.0f}, "
      f"loop overhead {np.mean(mm[s, 0][1:] - mm[s, 2][:-1]):.0f}, K blocks per round {np.mean(mm[s, 3]):.2f}")
print(f"producer K blocks: mean period {np.mean(np.diff(prod[20:-5, 0])):.0f}, wait {np.mean(prod[20:-5, 1] - prod[20:-5, 0]):.0f}")
