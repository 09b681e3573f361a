#!/usr/bin/env python
"""Per-chunk epilogue timeline (TLXCV_DEBUG_TRACE_CONV=<file> with TLXCV_DEBUG_ABLATE=64): warp 2 of CTA 0."""
import sys
import numpy as np

L = 4096
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(3, L).astype(np.int64)
e = a[2][a[2] > 0]
n = len(e) // 6
e = e[:6 * n].reshape(n, 6)
e = e - e[0, 0]
names = ["wait slot/res", "ldtm wait", "math", "stage+fence", "store+refill"]
print(f"chunks {n}")
for k in range(min(n, int(sys.argv[2]) if len(sys.argv) > 2 else 16)):
    print(k, e[k, 0], " ".join(f"{int(x):6d}" for x in np.diff(e[k])), "| gap to next", int(e[k + 1, 0] - e[k, 5]) if k + 1 < n else 0)
s = slice(4, n - 2)
d = np.diff(e[s], axis=1).mean(axis=0)
print("mean per chunk: " + ", ".join(f"{nm} {v:.0f}" for nm, v in zip(names, d)) + f"; total {np.diff(e[s][:, [0, 5]], axis=1).mean():.0f}; period {np.diff(e[s][:, 0]).mean():.0f}")
