import sys, os
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
from ppo_and_friends_b200 import _lib
_lib.LIB_PATH = os.path.join(os.getcwd(), "scratch", sys.argv[1])
import torch, bench
pk = bench.peaks()
res = bench.microbench_c2(pk, iters=5)
print(sys.argv[1], {k: round(p['frac_of_hbm_peak'], 3) for k, p in res["pieces"].items()}, round(res["total_frac_of_hbm_peak"], 3))
