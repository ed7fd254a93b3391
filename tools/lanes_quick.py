"""All-move sweep on lane-split walkers (placement 3), one line per ndim: steps/s (knob experiments: tools/sweep.sh)."""
import sys, os, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcintegratorplusplus_b200 as m
for nd in (64, 128):
    mci = m.MCI(nd); mci.setStatePlacement(3); mci.setRngMode(0); mci.setSeed(1337); mci.setNWalkers(65536); mci.setTrialMove(m.MoveType.All)
    mci.setX([0.1 if j % 2 == 0 else -0.05 for j in range(nd)]); mci.setMRT2Step(3.0/nd**0.5)
    mci.addSamplingFunction(m.ExpNDPDF(nd)); mci.addObservable(m.XND(nd), 20, 1)
    mci.integrate(400, False, False)
    best = 1e30
    for _ in range(3):
        avg, err = mci.integrate(2000, False, False)
        best = min(best, mci.timings()["walk_ms"])
    print(json.dumps({"ndim": nd, "steps_per_s": 65536*2000/(best*1e-3), "walk_ms": best, "acc": mci.getAcceptanceRate(), "avg0": float(avg[0])}), flush=True)
