# Per-launch times of the minibatch step INSIDE a CUDA graph: capture the chain cut after n launches
# (debug build, PPOAF_STOP_AFTER) and difference the replay times.  Usage: python scratch/prefix_times.py [c4] [ffma|tcgen05]
import sys, os, ctypes as C
import torch
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
from ppo_and_friends_b200 import _lib
_lib.LIB_PATH = os.path.join(os.getcwd(), "scratch", "libppoaf_timing.so")
import bench
from ppo_and_friends_b200 import ops
from ppo_and_friends_b200.ppo import _Loader, ppo_batch_train
from ppo_and_friends_b200._lib import load, check, stream_ptr
wl = sys.argv[1] if len(sys.argv) > 1 else "c4"
if len(sys.argv) > 2: ops.set_gemm_backend(sys.argv[2])
w = bench.WORKLOADS[wl]
ro, pol = bench.build_workload(w, 0, "cuda:0")
hp = bench.HotPath(w, ro, pol)
pol.initialize_dataset(); pol.dataset.ring = pol._ring; pol.dataset._seg = hp.seg; pol.finalize_dataset()
ds = pol.dataset
loader = _Loader(ds, w["B"])
ppo_batch_train(hp.state, loader, "pol")
eng = pol._engine
bufs = eng._bufs(ds, w["B"])
lib = load()
fused = os.environ.get("PPOAF_NO_HEAD_FUSION") is None
names = (["fwd0", "fwd1", "fwd2", "loss+heads", "bwd2+dW3", "bwd1", "bwd0", "adam"] if fused else
         ["fwd0", "fwd1", "fwd2", "fwd3", "loss", "bwd3", "bwd2", "bwd1", "bwd0", "adam"])
NK = len(names)
def graph_time(n, reps=20, inner=64):
    os.environ["PPOAF_STOP_AFTER"] = str(min(n, NK - 1))
    eng.mb_cursor.zero_()
    s = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(s):
        def body():
            check(lib.ppoaf_ppo_minibatch_grads(C.byref(eng.cfg), C.byref(bufs), stream_ptr()))
            if n >= NK: check(lib.ppoaf_ppo_minibatch_apply(C.byref(eng.cfg), C.byref(bufs), stream_ptr()))
        body(); torch.cuda.synchronize()
        with torch.cuda.graph(g, stream=s):
            for _ in range(inner): body()
        best = 1e9
        for _ in range(reps):
            eng.mb_cursor.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            torch.cuda.synchronize(); e0.record(s); g.replay(); e1.record(s); torch.cuda.synchronize()
            best = min(best, e0.elapsed_time(e1) * 1e3 / inner)
    return best
prev = 0.0
for n in range(1, NK + 1):
    t = graph_time(n)
    print(f"{names[n-1]:10s} cumulative {t:7.2f} us   delta {t - prev:6.2f} us", flush=True)
    prev = t
if hasattr(lib, "ppoaf_debug_loss_stamps") or True:
    try:
        lib.ppoaf_debug_loss_stamps.argtypes = [C.c_void_p]; lib.ppoaf_debug_loss_stamps.restype = C.c_int
        buf = (C.c_longlong * 16)(); lib.ppoaf_debug_loss_stamps(buf)
        print("loss kernel stamps (cycles from CTA entry; 0-7 CTA 0, 8-10 last CTA):", list(buf)[:11])
    except Exception as e:
        print("no loss stamps", e)
