#!/usr/bin/env python
"""Print the pipeline timeline dumped by TLXCV_DEBUG_TRACE_SLAB=<file> (CTA 0; cycles relative to the first event)."""
import sys
import numpy as np

L = 4096
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(3, L).astype(np.int64)
t0 = min(int(a[r][0]) for r in range(3) if a[r][0] > 0)
lo, hi = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (8, 20)
prod = a[0][a[0] > 0] - t0
mma = a[1][a[1] > 0] - t0
epi = a[2][a[2] > 0] - t0
print(f"events: producer {len(prod)}, mma {len(mma)}, epilogue {len(epi)}; total cycles {max(prod.max(), mma.max(), epi.max())}")
print("step | producer: wait_start slot_free issued | mma: start acc_free landed issued | epi: start acc_done staged copied")
for k in range(lo, hi):
    p = prod[3 * k:3 * k + 3] if 3 * k + 3 <= len(prod) else []
    m = mma[4 * k:4 * k + 4] if 4 * k + 4 <= len(mma) else []
    e = epi[4 * k:4 * k + 4] if 4 * k + 4 <= len(epi) else []
    print(k, "|", " ".join(f"{x:8d}" for x in p), "|", " ".join(f"{x:8d}" for x in m), "|", " ".join(f"{x:8d}" for x in e))
n = min(len(mma) // 4, len(epi) // 4)
if n > 12:
    d = lambda arr, i, j: np.mean(arr[j::4][8:n - 2] - arr[i::4][8:n - 2])
    print(f"mma per step: period {np.mean(np.diff(mma[0::4][8:n-2])):.0f}; wait acc {d(mma,0,1):.0f}, wait operands {d(mma,1,2):.0f}, issue {d(mma,2,3):.0f}")
    print(f"epi per step: period {np.mean(np.diff(epi[0::4][8:n-2])):.0f}; wait acc {d(epi,0,1):.0f}, ld+math+stage {d(epi,1,2):.0f}, bar+copy {d(epi,2,3):.0f}")
    np_ = min(len(prod) // 3, n)
    print(f"producer per group: period {np.mean(np.diff(prod[0::3][8:np_-2])):.0f}; wait empty {np.mean(prod[1::3][8:np_-2]-prod[0::3][8:np_-2]):.0f}, issue {np.mean(prod[2::3][8:np_-2]-prod[1::3][8:np_-2]):.0f}")
    # latency from mma issue of step k to epilogue acc_done of step k
    print(f"mma issued -> epilogue sees accumulator: {np.mean(epi[1::4][8:n-2] - mma[3::4][8:n-2]):.0f} cycles")
