"""Developer helper (not a pytest): GPU timeline of async fwd+bwd steps via torch.profiler (CUPTI):
per-kernel/memset/memcpy durations and the idle gaps between consecutive GPU activities."""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import gpu_profile as gp
import torch
from torch.profiler import profile, ProfilerActivity
P, W, H = 2_000_000, 1600, 1060
if len(sys.argv) > 3: P, W, H = map(int, sys.argv[1:4])
sc, tt, gct, got, rast, leaves, m2d = gp.setup(P, W, H)
for _ in range(5): gp.product_step(rast, leaves, m2d, gct, got)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(6): gp.product_step(rast, leaves, m2d, gct, got)
    torch.cuda.synchronize()
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "gpurun_out", "timeline.json")
prof.export_chrome_trace(out)
ev = [e for e in json.load(open(out))["traceEvents"] if e.get("cat") in ("kernel", "gpu_memset", "gpu_memcpy")]
ev.sort(key=lambda e: e["ts"])
# take the 4th step: find occurrences of surfel_preprocess_fwd
idx = [i for i, e in enumerate(ev) if "surfel_preprocess_fwd" in e["name"]]
a, b = idx[3], idx[4]
prev_end = ev[a - 1]["ts"] + ev[a - 1]["dur"]
busy = 0.0
for e in ev[a:b]:
    gap = e["ts"] - prev_end
    print(f"gap {gap:8.1f} us | {e['dur']:8.1f} us  {e['name'][:90]}")
    prev_end = e["ts"] + e["dur"]
    busy += e["dur"]
print("step span %.1f us, busy %.1f us" % (ev[b]["ts"] - ev[a]["ts"], busy))
