"""Single-index move sweep (the reference's bench_throughput_ndim_single shape with Block(20)), one line per ndim (knob experiments: tools/sweep.sh)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import mcintegratorplusplus_b200 as m
for nd in (8, 16, 32, 64):
    mci = bench.c3_mci(m, "vec", nd, 65536, None)
    mci.integrate(400, False, False)
    best = 1e30
    for _ in range(3):
        avg, err = mci.integrate(4000, False, False)
        best = min(best, mci.timings()["walk_ms"])
    print(json.dumps({"ndim": nd, "steps_per_s": 65536*4000/(best*1e-3), "walk_ms": best, "acc": mci.getAcceptanceRate(), "avg0": float(avg[0])}), flush=True)
