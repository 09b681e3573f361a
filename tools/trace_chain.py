#!/usr/bin/env python
"""Timeline of conv_chain_kernel CTA 0 dumped by TLXCV_DEBUG_TRACE_CHAIN=<file> (debug build)."""
import sys
import numpy as np

L = 4096
a = np.fromfile(sys.argv[1], dtype=np.uint64).reshape(3, L).astype(np.int64)
t0 = min(int(a[r][0]) for r in range(3) if a[r][0] > 0)
prod, mma, epi = [a[r][a[r] > 0] - t0 for r in range(3)]
nt = len(mma) // 6
n_chunks = int(sys.argv[3]) if len(sys.argv) > 3 else 2
print(f"tiles {nt}; total cycles {max(prod.max(), mma.max(), epi.max())}; epilogue events {len(epi)}")
print("tile | producer: start issued | mma: start acc1_free before_a2 a2_ready acc2_free end | epilogue warp 2 per chunk: start acc_done done")
for k in range(min(nt, int(sys.argv[2]) if len(sys.argv) > 2 else 10)):
    e = epi[3 * n_chunks * k:3 * n_chunks * (k + 1)]
    print(k, "|", " ".join(f"{x:7d}" for x in prod[2 * k:2 * k + 2]), "|", " ".join(f"{x:7d}" for x in mma[6 * k:6 * k + 6]), "|",
          " ".join(f"{x:7d}" for x in e))
if nt > 6:
    s = slice(2, nt - 1)
    m = mma[:6 * nt].reshape(nt, 6)[s]
    print(f"mma per tile: period {np.mean(np.diff(m[:, 0])):.0f}; wait acc1 {np.mean(m[:, 1] - m[:, 0]):.0f}; G1 part 1 {np.mean(m[:, 2] - m[:, 1]):.0f}; "
          f"wait A2 {np.mean(m[:, 3] - m[:, 2]):.0f}; wait acc2 {np.mean(m[:, 4] - m[:, 3]):.0f}; rest {np.mean(m[:, 5] - m[:, 4]):.0f}")
