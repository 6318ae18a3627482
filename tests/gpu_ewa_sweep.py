"""Developer helper: 3DGS 1M / plane cfg-4 timings (bench_extras workloads) for the library selected by GSR_B200_LIB,
plus the library's per-kernel event times."""
import ctypes, json, os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bench_extras as be
import gsr_b200
for name in ("gauss_1m", "cfg4_plane"):
    r = be.WORKLOADS[name][0]("ours")
    print(os.path.basename(gsr_b200._lib.LIB_PATH), name, f"{r['ms']:.3f} ms ({r['p10_ms']:.3f}..{r['p90_ms']:.3f})", flush=True)
