import sys, os, time
import numpy as np, torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import bench
from ppo_and_friends_b200.ppo import _get_engine, _Loader, ppo_batch_train
wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
w = bench.WORKLOADS[wl]
ro, pol = bench.build_workload(w, 0, "cuda:0")
hp = bench.HotPath(w, ro, pol)
def ev_time(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
# finalize
pol.initialize_dataset(); pol.dataset.ring = pol._ring; pol.dataset._seg = hp.seg
t0 = time.perf_counter(); pol.finalize_dataset(); torch.cuda.synchronize(); print("finalize wall ms", (time.perf_counter()-t0)*1e3)
def fin():
    pol.initialize_dataset(); pol.dataset.ring = pol._ring; pol.dataset._seg = hp.seg; pol.finalize_dataset()
print("finalize us (events)", ev_time(fin, 5))
ds = pol.dataset
loader = _Loader(ds, w["B"])
t0 = time.perf_counter(); ppo_batch_train(hp.state, loader, "pol"); torch.cuda.synchronize(); print("first epoch wall ms", (time.perf_counter()-t0)*1e3)
t0 = time.perf_counter(); ppo_batch_train(hp.state, loader, "pol"); torch.cuda.synchronize(); print("second epoch wall ms", (time.perf_counter()-t0)*1e3)
eng = pol._engine
g = list(eng._graphs.values())[0][0]
def replay():
    g.replay()
eng.mb_cursor.zero_()
print("graph replay us", ev_time(replay, 60)); eng.mb_cursor.zero_()
bufs = eng._bufs(ds, w["B"])
def eager():
    eng._step_eager(bufs)
print("eager step us", ev_time(eager, 60)); eng.mb_cursor.zero_()
t0 = time.perf_counter()
for _ in range(60): g.replay()
print("replay CPU launch us", (time.perf_counter()-t0)*1e6/60); torch.cuda.synchronize(); eng.mb_cursor.zero_()
import ctypes as C
from ppo_and_friends_b200._lib import load, check, stream_ptr
lib = load()
def grads_only():
    check(lib.ppoaf_ppo_minibatch_grads(C.byref(eng.cfg), C.byref(bufs), stream_ptr()))
def apply_only():
    check(lib.ppoaf_ppo_minibatch_apply(C.byref(eng.cfg), C.byref(bufs), stream_ptr()))
print("eager grads us", ev_time(grads_only, 60)); print("eager apply us", ev_time(apply_only, 30)); eng.mb_cursor.zero_()
# per-kernel-ish: forward only of each net through mlp_forward
from ppo_and_friends_b200 import ops
idx = eng._perm_dev[:w["B"]].contiguous()
print("actor fwd (4 kernels) us", ev_time(lambda: ops.mlp_forward(pol.actor.desc, pol.actor.flat, ds.observations, idx=idx), 50))
print("critic fwd (4 kernels) us", ev_time(lambda: ops.mlp_forward(pol.critic.desc, pol.critic.flat, ds.critic_observations, idx=idx), 50))
