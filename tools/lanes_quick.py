"""All-move sweep, lane-split walkers (placement 3) against the automatic placement, one line per (ndim, placement): python tools/lanes_quick.py [NDIM ...]"""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcintegratorplusplus_b200 as m
for nd in ([int(a) for a in sys.argv[1:]] or [64, 128]):
    for placement in (-1, 3):
        mci = m.MCI(nd); mci.setStatePlacement(placement); mci.setRngMode(0); mci.setSeed(1337); mci.setNWalkers(65536); mci.setTrialMove(m.MoveType.All)
        mci.setX([0.1 if j % 2 == 0 else -0.05 for j in range(nd)]); mci.setMRT2Step(3.0/nd**0.5)
        mci.addSamplingFunction(m.ExpNDPDF(nd)); mci.addObservable(m.XND(nd), 20, 1)
        try:
            mci.integrate(400, False, False)
        except Exception as e:
            print(json.dumps({"ndim": nd, "placement": placement, "error": str(e)[:100]}), flush=True)
            continue
        best = 1e30
        for _ in range(3):
            avg, err = mci.integrate(2000, False, False)
            best = min(best, mci.timings()["walk_ms"])
        print(json.dumps({"ndim": nd, "placement": placement, "steps_per_s": 65536*2000/(best*1e-3), "walk_ms": best, "acc": mci.getAcceptanceRate(), "avg0": float(avg[0])}), flush=True)
