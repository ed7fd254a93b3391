"""MultiStepMove sweep (BASELINE configs[2] shape: Gauss main pdf, ExpNDPDF sub-pdf, XND Block(20)), one line per ndim (knob experiments: tools/sweep.sh)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import mcintegratorplusplus_b200 as m
for nd in ([int(a) for a in sys.argv[1:]] or [16, 32, 64]):
    mci = bench.c3_mci(m, "multistep", nd, 65536, None)
    nmc = max(200, 8000//(2*nd)//20*20)
    mci.integrate(40, False, False)
    best = 1e30
    for _ in range(3):
        avg, err = mci.integrate(nmc, False, False)
        best = min(best, mci.timings()["walk_ms"])
    print(json.dumps({"ndim": nd, "outer_steps_per_s": 65536*nmc/(best*1e-3), "walk_ms": best, "acc": mci.getAcceptanceRate(), "avg0": float(avg[0])}), flush=True)
