"""Per-phase timing of the persistent step kernel (debug stamps): PPOAF_FUSED_STAMPS=1 python scratch/fused_stamps.py [c4|c3|c5|c1]"""
import ctypes as C, os, sys
os.environ["PPOAF_FUSED_STAMPS"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
import bench
from ppo_and_friends_b200 import _lib
from ppo_and_friends_b200.ppo import PPOUpdateState, _Loader, ppo_batch_train
from helpers import run_device_rollout

w = bench.WORKLOADS[sys.argv[1] if len(sys.argv) > 1 else "c4"]
dev = torch.device("cuda:0")
ro, pol = bench.build_workload(w, 0, dev)
ds = run_device_rollout(pol, ro)
state = PPOUpdateState({"pol": pol}, batch_size=w["B"], epochs_per_iter=1)
loader = _Loader(ds, w["B"])
for _ in range(3):
    ppo_batch_train(state, loader, "pol")
torch.cuda.synchronize()
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(5):
    ppo_batch_train(state, loader, "pol")
t1.record(); torch.cuda.synchronize()
n_mb = (len(ds) + w["B"] - 1) // w["B"]
print("us per minibatch step (incl. per-epoch host work):", t0.elapsed_time(t1) * 1e3 / 5 / n_mb, "n_mb", n_mb)
lib = _lib.load()
lib.ppoaf_debug_fused_stamps.restype = C.c_int
n = C.c_int32(0)
buf = (C.c_longlong * (3 * 32 * 4 + 2 * 32 * 32))()
rc = lib.ppoaf_debug_fused_stamps(buf, C.byref(n))
a = np.array(buf[:3 * n.value * 4], dtype=np.int64).reshape(3, n.value, 4)
names = {0: "cta0", 1: "mid", 2: "last"}
for s in range(3):
    rows = [r for r in a[s] if r[0] != 0]
    if not rows: continue
    base = rows[0][0]
    print(names[s], "phase: wait_clk  work_clk   (start at clk)")
    for i, r in enumerate(rows):
        print(f"  {i:2d}  barrier {r[1]-r[0]:7d}  work {r[2]-r[1]:7d}   @{r[0]-base:8d}")
    print("  step total clk:", rows[-1][2] - rows[0][0], " = us @1.965GHz:", (rows[-1][2] - rows[0][0]) / 1965.0)

nph = n.value
t = np.array(buf[3 * nph * 4: 3 * nph * 4 + nph * 32], dtype=np.int64).reshape(nph, 32)
for ph in range(nph):
    r = t[ph]
    if r[0] == 0: continue
    if r[28] == 0:
        print(f"phase {ph} detail (loss: entry,staged,loads,heads,actor,dX,partials | adam: entry,sqloads,fold,scalars,compute,-,bar):", [int(x - r[0]) for x in r[:7]])
        continue
    print(f"tile detail phase {ph}: plan+first loads {r[1]-r[0]}, mainloop {r[2]-r[1]}, acc wait {r[3]-r[2]}, epilogue {r[28]-r[3]}")
    print("   epilogue (first 8 columns): tmem loads", int(r[20]-r[3]), "sums", int(r[21]-r[20]), "rest (act, stores, 2nd column group)", int(r[28]-r[21]))
    print("   chunk 4 converter chain: raw ready", int(r[10]-r[0]), "lds+raw_free", int(r[29]-r[0]), "split", int(r[30]-r[0]), "mma_free ok", int(r[31]-r[0]), "sttm+wait", int(r[26]-r[0]), "proxy fence", int(r[27]-r[0]), "arrive", int(r[11]-r[0]))
    print("   chunk (after free-wait, after arrive):", [(int(r[4+2*c]-r[0]), int(r[5+2*c]-r[0])) for c in range(12) if r[4+2*c]])

m = np.array(buf[3 * nph * 4 + nph * 32: 3 * nph * 4 + 2 * nph * 32], dtype=np.int64).reshape(nph, 32)
for ph in range(nph):
    r = m[ph]
    if r[0] == 0: continue
    base = t[ph][0]
    print(f"mma thread phase {ph}: per chunk (enter wait, full ok, issued+commit) rel. to tile start:",
          [(int(r[3*c]-base), int(r[3*c+1]-base), int(r[3*c+2]-base)) for c in range(8) if r[3*c]])
    print(f"   producer thread 0, first 4 chunks (after raw_free wait, after issue+arrive):", [(int(r[24+2*c]-base), int(r[25+2*c]-base)) for c in range(4) if r[24+2*c]])
