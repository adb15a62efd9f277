import sys, os, json
sys.path.insert(0, os.getcwd()); sys.path.insert(0, "tests")
import torch, bench
pk = bench.peaks()
res = bench.microbench_c2(pk, iters=int(sys.argv[1]) if len(sys.argv) > 1 else 5)
for k, p in res["pieces"].items():
    print(f"{k:22s} {p['ms']*1e3:9.1f} us  {p['gbs']:8.1f} GB/s  frac {p['frac_of_hbm_peak']:.3f}")
print("total", res["total_ms"], "ms", res["total_gbs"], "GB/s frac", res["total_frac_of_hbm_peak"])
