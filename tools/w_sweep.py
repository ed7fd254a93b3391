"""Cliff finder: throughput against the number of walkers for one kernel family each (register path, lane-split, shared-memory, MultiStepMove, global
memory). A launch rule that goes wrong at some job size shows up as a drop of steps/s below its smaller neighbour. python tools/w_sweep.py [prebuild]"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import bench  # noqa: E402
import mcintegratorplusplus_b200 as m  # noqa: E402

prebuild = len(sys.argv) > 1 and sys.argv[1] == "prebuild"
FAMILIES = [("all", 3, 20000), ("all", 8, 20000), ("all", 32, 4000), ("vec", 64, 8000), ("multistep", 8, 2000), ("multistep", 32, 600), ("all", 256, 400)]
WS = [1, 33, 1000, 4736, 9472, 18944, 30000, 37888, 50000, 75776, 100000, 151552, 200000]
for move, nd, nmc in FAMILIES:
    prev = 0.
    for W in WS:
        mci = bench.c3_mci(m, move, nd, W, None)
        if prebuild:
            mci.prebuild()
            continue
        mci.integrate(nmc, False, False)
        mci.integrate(nmc, False, False)
        t = mci.timings()
        rate = W*nmc/(t["walk_ms"]*1e-3)
        flag = "  <-- CLIFF" if rate < 0.85*prev else ""
        print("%-9s ndim %3d W %6d nmc %5d walk %8.3f ms total %8.3f ms %.3e steps/s%s" % (move, nd, W, nmc, t["walk_ms"], t["total_ms"], rate, flag), flush=True)
        prev = rate
