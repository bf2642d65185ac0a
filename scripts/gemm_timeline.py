"""Summarise a ROREG_DEBUG_GEMM_TRACE file (clock64 timeline of CTA 0 of one gather-GEMM launch, trace build only)."""
import sys
import numpy as np
for fn in sys.argv[1:]:
    hdr = open(fn).readline().strip()
    t = np.loadtxt(fn, dtype=np.int64)
    n = int((t[:, 2] > 0).sum()); t0 = t[0, 6] if t[0, 6] > 0 else t[0, 0]
    print(fn, hdr, "k-chunk iterations traced", n)
    e = t[:n].astype(np.float64) - t0
    per = np.diff(e[:, 2])
    print(f"  stage period (MMA warp sees full -> next): median {np.median(per):.0f} mean {per.mean():.0f} p10 {np.percentile(per, 10):.0f} p90 {np.percentile(per, 90):.0f} clk")
    print(f"  loader: empty seen -> copies issued: median {np.median(e[:, 1] - e[:, 0]):.0f};  issued -> MMA warp sees full: median {np.median(e[:, 2] - e[:, 1]):.0f};"
          f"  full seen -> commit issued: {np.median(e[:, 3] - e[:, 2]):.0f}")
    lat = e[4:n, 0] - e[:n - 4, 3]
    print(f"  commit(it) -> loader sees empty(it+4): median {np.median(lat):.0f} p10 {np.percentile(lat, 10):.0f} p90 {np.percentile(lat, 90):.0f}")
    nt = int((t[:, 4] > 0).sum())
    print("  tiles:", nt, "| epilogue start", (t[:nt, 4] - t0), "end", (t[:nt, 5] - t0), "| MMA warp got the accumulator", (t[:nt, 6] - t0))
