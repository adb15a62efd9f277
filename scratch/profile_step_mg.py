import sys, os, time
import numpy as np, torch, torch.distributed as dist
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
if os.environ.get("PPOAF_TIMING_LIB"):
    from ppo_and_friends_b200 import _lib as _l
    _l.LIB_PATH = os.path.join(os.getcwd(), "scratch", "libppoaf_timing.so")
import bench
rank, world, local = bench.dist_env()
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
from ppo_and_friends_b200.ppo import _Loader, ppo_batch_train
w = bench.WORKLOADS["c4"]
ro, pol = bench.build_workload(w, rank, f"cuda:{local}")
hp = bench.HotPath(w, ro, pol)
pol.initialize_dataset(); pol.dataset.ring = pol._ring; pol.dataset._seg = hp.seg; pol.finalize_dataset()
ds = pol.dataset
loader = _Loader(ds, w["B"])
ppo_batch_train(hp.state, loader, "pol"); ppo_batch_train(hp.state, loader, "pol")
eng = pol._engine
def ev_time(fn, n):
    torch.cuda.synchronize(); dist.barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / n
def full_step():
    eng._launch_step(ds, w["B"])
eng.mb_cursor.zero_()
t_full = ev_time(full_step, 60); eng.mb_cursor.zero_()
bufs = eng._bufs(ds, w["B"], 0)
g = eng._capture(lambda: eng._grads(bufs))
t_grads = ev_time(lambda: g.replay(), 60)
# peer kernel alone, in lockstep (both ranks call it the same number of times)
from ppo_and_friends_b200._lib import stream_ptr
def peer_only():
    eng.peer.allreduce_adam(0, pol.nets, eng.mb_cursor, eng.hparams, stream_ptr())
t_peer = ev_time(peer_only, 60) if eng.peer is not None else float("nan")
print(f"[rank {rank}] full step {t_full:.1f} us | grads graph {t_grads:.1f} us | peer allreduce+adam eager {t_peer:.1f} us", flush=True)
if os.environ.get("PPOAF_TIMING_LIB"):
    import ctypes as C
    from ppo_and_friends_b200 import _lib
    lib = _lib.load()
    lib.ppoaf_debug_peer_stamps.argtypes = [C.c_void_p]
    buf = (C.c_longlong * 16)(); lib.ppoaf_debug_peer_stamps(buf)
    st = list(buf)[:8]
    print(f"[rank {rank}] peer kernel stamps (cycles):", [x - st[0] for x in st], flush=True)
dist.barrier(); dist.destroy_process_group()
