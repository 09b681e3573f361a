#!/usr/bin/env python
"""Tile-boundary timeline of one epilogue warp (TLXCV_FINE_TRACE build, TLXCV_DEBUG_ABLATE=320, TLXCV_DEBUG_TRACE_CONV=<file>)."""
import sys
import numpy as np

L = 4096
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(3, L).astype(np.int64)
e = a[2][a[2] > 0]
n = len(e) // 5
e = e[:5 * n].reshape(n, 5)
e = e - e[0, 0]
print("tile: start | set-up, wait accumulator, stage scale/shift, chunk loop | gap to next tile start")
for k in range(min(n, 12)):
    print(k, e[k, 0], "|", " ".join(f"{int(x):6d}" for x in np.diff(e[k])), "|", int(e[k + 1, 0] - e[k, 4]) if k + 1 < n else 0)
s = slice(2, n - 1)
print("mean:", np.diff(e[s], axis=1).mean(axis=0).round(0), "gap", (e[3:n, 0] - e[2:n - 1, 4]).mean().round(0))
