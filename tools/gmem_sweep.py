"""Global-memory walker placement: block size x walker count sweep (vec move, ndim 256 / 1024). Run on a B200 via gpurun."""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mcintegratorplusplus_b200 as m  # noqa: E402


def run(nd, W, bs, nmc, move="vec", accu=20):
    mci = m.MCI(nd)
    mci.setRngMode(0)
    mci.setSeed(1337)
    mci.setNWalkers(W)
    if bs:
        mci.setBlockSize(bs)
    mci.setTrialMove(m.MoveType.Vec if move == "vec" else m.MoveType.All)
    mci.setX([0.1 if j % 2 == 0 else -0.05 for j in range(nd)])
    mci.setMRT2Step(3.0 if move == "vec" else 3.0/nd**0.5)
    mci.addSamplingFunction(m.ExpNDPDF(nd))
    if accu >= 0:
        mci.addObservable(m.XND(nd), accu, 1)
    else:
        mci.addObservable(m.X2Sum(nd), 0, 1)
    mci.integrate(200, False, False)
    mci.integrate(nmc, False, False)
    t = mci.timings()
    print(json.dumps({"ndim": nd, "move": move, "walkers": W, "block": bs, "accu": accu, "nmc": nmc, "steps_per_s": W*nmc/(t["walk_ms"]*1e-3), "walk_ms": t["walk_ms"]}), flush=True)


if __name__ == "__main__":
    for W in (16384, 65536):
        for bs in (32, 64, 128, 256):
            run(256, W, bs, 4000)
    run(256, 65536, 128, 4000, accu=-1)   # scalar observable: the step itself, without the O(ndim) accumulation
    run(1024, 65536, 64, 2000)
    run(1024, 65536, 64, 2000, accu=-1)
    run(256, 65536, 64, 500, move="all")
    run(1024, 32768, 64, 200, move="all")
