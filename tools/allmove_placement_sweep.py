"""All-moves at ndim 96 .. 256: shared-memory vs global-memory walker placement and block size (after the proto-value arrays left the footprint)."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mcintegratorplusplus_b200 as m  # noqa: E402


def run(nd, placement, bs, nmc, W=65536, accu=20):
    mci = m.MCI(nd)
    mci.setRngMode(0)
    mci.setSeed(1337)
    mci.setNWalkers(W)
    mci.setStatePlacement(placement)
    if bs:
        mci.setBlockSize(bs)
    mci.setTrialMove(m.MoveType.All)
    mci.setX([0.1 if j % 2 == 0 else -0.05 for j in range(nd)])
    mci.setMRT2Step(3.0/nd**0.5)
    mci.addSamplingFunction(m.ExpNDPDF(nd))
    mci.addObservable(m.XND(nd), accu, 1)
    try:
        mci.integrate(200, False, False)
        mci.integrate(nmc, False, False)
    except Exception as e:  # a placement that does not fit fails loudly: record and go on
        print(json.dumps({"ndim": nd, "placement": placement, "block": bs, "error": str(e)[:100]}), flush=True)
        return
    t = mci.timings()
    print(json.dumps({"ndim": nd, "placement": placement, "block": bs, "accu": accu, "steps_per_s": W*nmc/(t["walk_ms"]*1e-3), "walk_ms": t["walk_ms"]}), flush=True)


if __name__ == "__main__":
    for nd, nmc in ((64, 4000), (96, 2000), (128, 1000), (192, 1000), (256, 500)):
        run(nd, -1, 0, nmc)         # automatic (0 would force the register placement)
        for bs in (32, 64, 128, 256):
            run(nd, 2, bs, nmc)     # global memory
