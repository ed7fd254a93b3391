"""Headline workload (C2: 3-D Gaussian + x^2, all-move, Simple accumulator, 65536 walkers x 1e5 steps), best of 3 walk-kernel times (knob experiments: tools/sweep.sh)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
import mcintegratorplusplus_b200 as m
W = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
nmc = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
mci = g._bench_mci(m, nwalkers=W)
mci.integrate(2000, False, False)
best = 1e30
for _ in range(3):
    avg, err = mci.integrate(nmc, False, False)
    best = min(best, mci.timings()["walk_ms"])
print(json.dumps({"walkers": W, "samples_per_s": W*nmc/(best*1e-3), "walk_ms": best, "acc": mci.getAcceptanceRate(), "avg0": float(avg[0])}), flush=True)
