import sys, os, ctypes as C, torch
sys.path.insert(0, os.getcwd())
from ppo_and_friends_b200 import _lib
_lib.LIB_PATH = os.path.join(os.getcwd(), "scratch", "libppoaf_timing.so")
from ppo_and_friends_b200 import ops
ops.runtime_init()
lib = _lib.load()
lib.ppoaf_debug_gemm_stamps.argtypes = [C.c_void_p]; lib.ppoaf_debug_gemm_stamps.restype = C.c_int
for dims, rows in (([256, 256], 512), ([376, 256], 512), ([16, 64], 512)):
    desc = _lib.MlpDesc.make(dims, "tanh")
    offs, total = _lib.param_layout(desc)
    params = torch.randn(total, device="cuda") * 0.05
    xin = torch.randn(rows, dims[0], device="cuda"); out = torch.empty(rows, dims[1], device="cuda")
    for _ in range(5): ops.mlp_forward(desc, params, xin, out=out)
    torch.cuda.synchronize()
    buf = (C.c_longlong * 16)()
    lib.ppoaf_debug_gemm_stamps(buf)
    st = list(buf)[:15]
    print(dims, "cycles from kernel entry:", [s - st[0] for s in st])
