"""One MultiStepMove configuration of tools/ms_quick.py: python tools/ms_one.py NDIM (MCIG_CUBIN_OVERRIDE experiments: one kernel per process)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import mcintegratorplusplus_b200 as m
nd = int(sys.argv[1])
mci = bench.c3_mci(m, "multistep", nd, 65536, None)
nmc = max(200, 8000//(2*nd)//20*20)
mci.integrate(40, False, False)
best = 1e30
for _ in range(3):
    avg, err = mci.integrate(nmc, False, False)
    best = min(best, mci.timings()["walk_ms"])
print(json.dumps({"ndim": nd, "outer_steps_per_s": 65536*nmc/(best*1e-3), "walk_ms": best, "acc": mci.getAcceptanceRate(), "avg0": float(avg[0])}), flush=True)
