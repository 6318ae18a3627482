"""Developer helper (not a pytest): where a GS-SR-style 2DGS iteration spends GPU time (torch.profiler)."""
import os, sys, collections
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import harness  # noqa
import torch
from torch.profiler import profile, ProfilerActivity
from train_harness import MiniTwoDGSTrainer, MiniScaffold2DGSTrainer, MiniPGSRTrainer
flow = sys.argv[4] if len(sys.argv) > 4 else "2dgs"
P, W, H = (int(x) for x in (sys.argv[1:4] or (2_000_000, 1600, 1060)))
if flow == "scaffold":
    tr = MiniScaffold2DGSTrainer(P, W=W, H=H, impl="ours")
elif flow == "pgsr":
    tr = MiniPGSRTrainer(P, W=W, H=H, impl="ours")
else:
    tr = MiniTwoDGSTrainer(P, W, H, impl="ours", lambda_dist=0.0)
for _ in range(5): tr.step()
torch.cuda.synchronize()
N = 5
with profile(activities=[ProfilerActivity.CPU, ProfilerActivity.CUDA]) as prof:
    for _ in range(N): tr.step()
    torch.cuda.synchronize()
ev = [e for e in prof.events() if e.device_type == torch.autograd.DeviceType.CUDA]
tot = collections.Counter(); cnt = collections.Counter()
for e in ev:
    tot[e.name[:70]] += e.device_time_total if hasattr(e, "device_time_total") else e.cuda_time_total
    cnt[e.name[:70]] += 1
total = sum(tot.values())
print(f"GPU busy per iteration: {total/N/1e3:.2f} ms over {sum(cnt.values())/N:.0f} kernels/memops")
for k, v in tot.most_common(28):
    print(f"{v/N:9.1f} us  x{cnt[k]/N:5.1f}  {k}")
