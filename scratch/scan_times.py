import sys, os, ctypes as C
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
from ppo_and_friends_b200 import _lib
_lib.LIB_PATH = os.path.join(os.getcwd(), "scratch", "libppoaf_timing.so")
import numpy as np, torch
from ppo_and_friends_b200 import ops
ops.runtime_init()
n = 1 << 22
g = torch.Generator(device="cuda").manual_seed(1)
r = torch.randn(n, device="cuda", generator=g); v = torch.randn(n, device="cuda", generator=g)
ends = (torch.rand(n, device="cuda", generator=g) < 0.02)
ends[63::64] = True; ends[-1] = True
flag = ends.to(torch.uint8)
end_pos = torch.nonzero(ends).flatten()
n_seg = end_pos.numel()
off = torch.cat([torch.zeros(1, dtype=torch.int64, device="cuda"), end_pos + 1])
vb = torch.randn(n_seg, device="cuda"); rb = torch.randn(n_seg, device="cuda")
adv = torch.empty(n, device="cuda"); rtg = torch.empty(n, device="cuda")
flush = torch.zeros(64 << 20, device="cuda")
for _ in range(3):
    flush.add_(1.0)
    ops.gae_rtg_segscan(r, v, flag, off, vb, rb, 0.99, 0.95, True, adv, rtg)
torch.cuda.synchronize()
lib = _lib.load()
lib.ppoaf_debug_scan_times.argtypes = [C.c_void_p, C.c_int]
nt = 4096
buf = (C.c_ulonglong * (4 * nt))()
lib.ppoaf_debug_scan_times(buf, nt)
t = np.array(buf, dtype=np.float64).reshape(nt, 4)
t0 = t[:, 0].min()
print("kernel span us", (t[:, 3].max() - t0) / 1e3)
d = (t - t0) / 1e3
for q in (0, 100, 1000, 1183, 1184, 2000, 3000, 4095):
    print("cta", q, "start %.2f  agg %.2f  carry %.2f  end %.2f" % tuple(d[q]))
print("mean per-CTA duration us %.2f ; load+localscan %.2f ; lookback %.2f ; apply+store %.2f" % ((d[:,3]-d[:,0]).mean(), (d[:,1]-d[:,0]).mean(), (d[:,2]-d[:,1]).mean(), (d[:,3]-d[:,2]).mean()))
