#!/usr/bin/env python
"""Timeline of conv_tcgen05 CTA 0 dumped by TLXCV_DEBUG_TRACE_CONV=<file> (last conv launch)."""
import sys
import numpy as np

L = 4096
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(3, L).astype(np.int64)
t0 = min(int(a[r][0]) for r in range(3) if a[r][0] > 0)
prod, mma, epi = [a[r][a[r] > 0] - t0 for r in range(3)]
nt = len(mma) // 4
print(f"tiles {nt}; total cycles {max(prod.max(), mma.max(), epi.max())}")
print("tile | producer: start issued | mma: start acc_free operands issued | epi: start acc_done done")
for k in range(min(nt, int(sys.argv[2]) if len(sys.argv) > 2 else 12)):
    print(k, "|", " ".join(f"{x:8d}" for x in prod[2 * k:2 * k + 2]), "|", " ".join(f"{x:8d}" for x in mma[4 * k:4 * k + 4]), "|",
          " ".join(f"{x:8d}" for x in epi[3 * k:3 * k + 3]))
if nt > 4:
    s = slice(1, nt - 1)
    print(f"mma: period {np.mean(np.diff(mma[0::4][s])):.0f}; wait acc {np.mean(mma[1::4][s]-mma[0::4][s]):.0f}; wait first operands "
          f"{np.mean(mma[2::4][s]-mma[1::4][s]):.0f}; issue {np.mean(mma[3::4][s]-mma[2::4][s]):.0f}")
    print(f"epi: period {np.mean(np.diff(epi[0::3][s])):.0f}; wait acc {np.mean(epi[1::3][s]-epi[0::3][s]):.0f}; chunks {np.mean(epi[2::3][s]-epi[1::3][s]):.0f}")
    print(f"producer: period {np.mean(np.diff(prod[0::2][s])):.0f}; issue span {np.mean(prod[1::2][s]-prod[0::2][s]):.0f}")
