"""Time the pieces of the value-normalise microbench step (2^22 returns): python scratch/value_norm_pieces.py"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ppo_and_friends_b200 import ops
from ppo_and_friends_b200.utils.stats import RunningMeanStd
n = 1 << 22
rtg = torch.randn(n, device="cuda"); out = torch.empty_like(rtg)
vrms = RunningMeanStd(shape=()); tri = torch.empty(3, dtype=torch.float64, device="cuda")
flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
def timed(fn, iters=5):
    fn(); torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g): fn()
    g.replay(); torch.cuda.synchronize(); ts = []
    for _ in range(iters):
        flush.add_(1); torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); g.replay(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e3)
    return round(float(np.mean(ts)), 1)
print("batch_moments(dim=1) us:", timed(lambda: ops.batch_moments(rtg, 1, tri)))
print("stats_merge us:", timed(lambda: ops.stats_merge(vrms.state, tri, 1)))
print("normalize_clip(dim=1) us:", timed(lambda: ops.normalize_clip(rtg, vrms.state, 1, 1e-8, 1.0, -1.0, out)))
print("copy_ (torch) us:", timed(lambda: out.copy_(rtg)))
