import sys, os, json, numpy as np
sys.path.insert(0, '/root/repo')
import mcintegratorplusplus_b200 as m
def run(nd, bs_accu, move="all"):
    mci = m.MCI(nd); mci.setRngMode(0); mci.setSeed(1337); mci.setNWalkers(65536)
    mci.setTrialMove(m.MoveType.All)
    mci.setX([0.1 if j % 2 == 0 else -0.05 for j in range(nd)])
    mci.setMRT2Step(3.0/np.sqrt(nd))
    mci.addSamplingFunction(m.ExpNDPDF(nd)); mci.addObservable(m.XND(nd), bs_accu, 1)
    mci.integrate(2000, False, False); mci.integrate(10000, False, False)
    t = mci.timings()
    print(os.environ.get("MCIG_SUMS_STATE_MIN"), nd, bs_accu, "%.3e" % (65536*10000/(t["walk_ms"]*1e-3)), flush=True)
for nd in (20, 24, 32, 48, 64):
    for b in (0, 20):
        run(nd, b)
