"""Run ONE bench-secondary C3 configuration a few times (for ncu captures): python tools/one_config.py MOVE NDIM [NMC] [PLACEMENT]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench
import mcintegratorplusplus_b200 as m
move, nd = sys.argv[1], int(sys.argv[2])
nmc = int(sys.argv[3]) if len(sys.argv) > 3 else 200
mci = bench.c3_mci(m, move, nd, 65536, None)
if len(sys.argv) > 4:
    mci.setStatePlacement(int(sys.argv[4]))
for _ in range(3):
    avg, err = mci.integrate(nmc, False, False)
    t = mci.timings()
    print(json.dumps({"move": move, "ndim": nd, "steps_per_s": 65536*nmc/(t["walk_ms"]*1e-3), "walk_ms": t["walk_ms"]}), flush=True)
