"""MultiStepMove at ndim 16..64: shared- vs global-memory walker placement x block size (C3). Run on a B200 via gpurun."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mcintegratorplusplus_b200 as m  # noqa: E402


def run(nd, placement, bs, nmc, W=65536):
    mci = m.MCI(nd)
    mci.setRngMode(0)
    mci.setSeed(1337)
    mci.setNWalkers(W)
    mci.setStatePlacement(placement)
    if bs:
        mci.setBlockSize(bs)
    mci.setTrialMove(m.MoveType.MultiStep, 1, sub_pdfs=[m.ExpNDPDF(nd)])
    mci.setX([0.1 if j % 2 == 0 else -0.05 for j in range(nd)])
    mci.setMRT2Step(0.5)
    mci.addSamplingFunction(m.Gauss(nd))
    mci.addObservable(m.XND(nd), 20, 1)
    try:
        mci.integrate(100, False, False)
        mci.integrate(nmc, False, False)
    except Exception as e:  # noqa: BLE001
        print(json.dumps({"ndim": nd, "placement": placement, "block": bs, "error": str(e)[:120]}), flush=True)
        return
    t = mci.timings()
    print(json.dumps({"ndim": nd, "placement": placement, "block": bs, "nmc": nmc, "steps_per_s": W*nmc/(t["walk_ms"]*1e-3), "walk_ms": t["walk_ms"]}), flush=True)


if __name__ == "__main__":
    for nd, nmc in ((16, 4000), (32, 2000), (64, 600)):
        for placement in (0, 2):  # 0 automatic (shared memory here), 2 global memory
            for bs in (0, 32, 64, 128, 256):
                run(nd, placement, bs, nmc)
