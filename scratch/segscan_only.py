"""GAE/RTG segmented scan alone at the M-C2 size (2^22 timesteps): python scratch/segscan_only.py [iters]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from ppo_and_friends_b200 import ops
n = 1 << 22
rng = np.random.default_rng(1234)
lens, left, gi = [], n, 0
geo = rng.geometric(0.02, size=n // 16)
while left > 0:
    L = int(min(left, 64, geo[gi])); gi += 1
    lens.append(L); left -= L
lens = np.asarray(lens, dtype=np.int64); n_seg = len(lens)
off = np.zeros(n_seg + 1, dtype=np.int64); np.cumsum(lens, out=off[1:])
term = rng.random(n_seg) < 0.5
flag = np.zeros(n, dtype=np.uint8); flag[off[1:] - 1] = 1 + 2 * term.astype(np.uint8)
d = lambda x: torch.as_tensor(x).cuda()
r = torch.randn(n, device="cuda"); v = torch.randn(n, device="cuda")
vb = torch.where(d(term), torch.zeros(n_seg, device="cuda"), torch.randn(n_seg, device="cuda")); rb = vb.clamp(-100, 100)
flag_d, off_d = d(flag), d(off)
adv = torch.empty(n, device="cuda"); rtg = torch.empty(n, device="cuda")
flush = torch.zeros(64 << 20, dtype=torch.float32, device="cuda")
iters = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ts = []
for i in range(iters + 2):
    flush.add_(1); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); ops.gae_rtg_segscan(r, v, flag_d, off_d, vb, rb, 0.99, 0.95, True, adv, rtg); e1.record(); torch.cuda.synchronize()
    if i >= 2: ts.append(e0.elapsed_time(e1) * 1e3)
print("segscan us (eager, incl. workspace memset):", [round(t, 1) for t in ts], "GB/s", (17 * n + 9 * n_seg) / np.mean(ts) / 1e3)
