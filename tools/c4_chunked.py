"""Chunked staging experiment: 65536 chains x 2^k Full + MJBlocker with the series forced through the chunked path (MCIG_CHUNK_BYTES)."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcintegratorplusplus_b200 as m  # noqa: E402


def run(W, k, budget, reps=2):
    if budget:
        os.environ["MCIG_CHUNK_BYTES"] = str(budget)
    else:
        os.environ.pop("MCIG_CHUNK_BYTES", None)
    mci = m.MCI(3)
    mci.setRngMode(0)
    mci.setSeed(1337)
    mci.setNWalkers(W)
    mci.addSamplingFunction(m.ThreeDimGaussianPDF())
    mci.addObservable(m.XSquared(), 1, 1, False, m.EstimatorType.MJBlocker)
    mci.setMRT2Step(1.0)
    for _ in range(reps):
        t0 = time.perf_counter()
        avg, err = mci.integrate(1 << k, False, False)
        wall = time.perf_counter() - t0
    t = mci.timings()
    print(json.dumps({"W": W, "k": k, "budget_MiB": (budget or 0) >> 20, "chunks": mci.getStagingChunks(), "total_ms": t["total_ms"], "walk_ms": t["walk_ms"],
                      "estim_ms": t["estim_ms"], "wall_ms": 1e3*wall, "samples_per_s": W*(1 << k)/(t["total_ms"]*1e-3), "avg": float(avg[0]), "err": float(err[0])}), flush=True)


if __name__ == "__main__":
    k = int(sys.argv[1]) if len(sys.argv) > 1 else 17
    run(65536, k, None)
    for mib in (32768, 8192, 2048, 512):
        run(65536, k, mib << 20)
