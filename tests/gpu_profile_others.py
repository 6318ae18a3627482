"""Developer helper (not a pytest): one realistic-size call of the kernels that bench.py does not exercise (plane rasterizer,
SH kernels, fused SSIM, fused post-processing) -- wrap with ncu for profiles/."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import numpy as np, torch
import harness as hz, synth
from train_harness import MiniTwoDGSTrainer, MiniPGSRTrainer
t = MiniTwoDGSTrainer(2_000_000, 1600, 1060, impl="ours", lambda_dist=0.0)
t.fused_ssim = t.fused_post = True
t.step()
p = MiniPGSRTrainer(1_000_000, W=1600, H=900, impl="ours")
p.fused_ssim = True
p.step()
torch.cuda.synchronize()
