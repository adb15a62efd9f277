import sys, os, torch, time
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
from ppo_and_friends_b200 import ops, _lib
ops.runtime_init()
def bench_graph(fn, n=200, reps=5):
    g = torch.cuda.CUDAGraph()
    torch.cuda.synchronize()
    with torch.cuda.graph(g):
        for _ in range(n): fn()
    g.replay(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps): g.replay()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / (n * reps)
x = torch.zeros(32, device="cuda")
print("tiny torch kernel chain us/kernel:", bench_graph(lambda: x.add_(1)))
big = torch.zeros(1 << 20, device="cuda")
print("4MB elementwise chain us/kernel:", bench_graph(lambda: big.add_(1)))
# one forward layer 512x256x256 as a 1-layer MLP (grid 64 tiles) and 2 nets-equivalent via N=512 (128 tiles)
for dims, rows in (([256, 256], 512), ([256, 512], 512), ([376, 512], 512), ([256, 1024], 512), ([256, 64], 512)):
    desc = _lib.MlpDesc.make(dims, "tanh")
    offs, total = _lib.param_layout(desc)
    params = torch.randn(total, device="cuda") * 0.05
    xin = torch.randn(rows, dims[0], device="cuda")
    out = torch.empty(rows, dims[1], device="cuda")
    t = bench_graph(lambda: ops.mlp_forward(desc, params, xin, out=out))
    fl = 2 * rows * dims[0] * dims[1]
    print(f"fwd layer {rows}x{dims[0]}x{dims[1]} tiles={(rows//32)*((dims[1]+63)//64)} us/kernel: {t:.2f}  ({fl/t/1e6:.1f} TFLOP/s)")
