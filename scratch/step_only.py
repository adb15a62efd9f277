"""Device time of the minibatch steps alone (no per-epoch host work): PPOAF_STEP=fused|chain python scratch/step_only.py [c4]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import torch
import bench
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
eng = pol._engine
n_full = len(ds) // w["B"]
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
reps = 10
def once():
    eng.mb_cursor.zero_()
    if eng.fused:
        eng._fused_steps(ds, w["B"], n_full)
    else:
        eng._launch_full_minibatches(ds, n_full)
once(); torch.cuda.synchronize()
ev[0].record()
for _ in range(reps):
    once()
ev[1].record(); torch.cuda.synchronize()
print("mode", "fused" if eng.fused else "chain", "us per minibatch step (device, steps only):", ev[0].elapsed_time(ev[1]) * 1e3 / reps / n_full, "n_full", n_full)
t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
t0.record()
for _ in range(5):
    ppo_batch_train(state, loader, "pol")
t1.record(); torch.cuda.synchronize()
print("   with per-epoch host work:", t0.elapsed_time(t1) * 1e3 / 5 / n_full)
