"""Secondary workloads of BASELINE.json (configs[2..4], pinned in SURVEY.md §8(d) C3/C4/C5) on one GPU: one JSON line each.
Not the driver's bench (that is bench.py, C2); these numbers go to profiles/ and DESIGN.md."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

import mcintegratorplusplus_b200 as m  # noqa: E402

HBM_GBS = 6550.7  # MEASURED_PEAKS.json (driver-measured copy bandwidth on this pool)


def c3_ndim(move, ndims=(1, 2, 4, 8, 16, 32, 64), W=65536, nmc=20000, placement=None):
    for nd in ndims:
        mci = m.MCI(nd)
        if placement is not None:
            mci.setStatePlacement(placement)
        mci.setRngMode(0)
        mci.setSeed(1337)
        mci.setNWalkers(W)
        if move == "vec":
            mci.setTrialMove(m.MoveType.Vec)
        elif move == "multistep":
            mci.setTrialMove(m.MoveType.MultiStep, 1, sub_pdfs=[m.ExpNDPDF(nd)] if nd > 1 else [])
        else:
            mci.setTrialMove(m.MoveType.All)
        mci.setX([0.1 if j % 2 == 0 else -0.05 for j in range(nd)])
        mci.setMRT2Step(3.0 if move == "vec" else (0.5 if move == "multistep" else 3.0/np.sqrt(nd)))
        mci.addSamplingFunction(m.ExpNDPDF(nd) if move != "multistep" else m.Gauss(nd))
        mci.addObservable(m.XND(nd), 20, 1)  # BlockAccumulator(20) + Uncorrelated
        mci.integrate(2000, False, False)
        t0 = time.perf_counter()
        avg, err = mci.integrate(nmc, False, False)
        wall = time.perf_counter() - t0
        t = mci.timings()
        print(json.dumps({"config": "C3_ndim_" + move, "placement": placement, "ndim": nd, "walkers": W, "nmc": nmc, "steps_per_s": W*nmc/(t["walk_ms"]*1e-3),
                          "walk_ms": t["walk_ms"], "estim_ms": t["estim_ms"], "wall_ms": 1e3*wall, "acceptance": mci.getAcceptanceRate(),
                          "max_abs_avg": float(np.max(np.abs(avg))), "max_err": float(np.max(err))}), flush=True)


def c4_estimators(W=65536, npow=14):
    nmc = 1 << npow
    for label, est in (("MJBlocker", m.EstimatorType.MJBlocker), ("Uncorrelated", m.EstimatorType.Uncorrelated), ("FCBlocker", m.EstimatorType.FCBlocker)):
        mci = m.MCI(3)
        mci.setRngMode(0)
        mci.setSeed(1337)
        mci.setNWalkers(W)
        mci.addSamplingFunction(m.ThreeDimGaussianPDF())
        mci.addObservable(m.XSquared(), 1, 1, False, est)  # FullAccumulator
        mci.setMRT2Step(1.0)
        mci.integrate(nmc, False, False)
        ts = []
        for _ in range(7):  # best of 7: single runs of a 1.5 ms kernel scatter by several percent
            avg, err = mci.integrate(nmc, False, False)
            ts.append(mci.timings())
        t = min(ts, key=lambda q: q["estim_ms"])
        t["walk_ms"] = min(q["walk_ms"] for q in ts)
        nbytes = 8.0*nmc*W
        print(json.dumps({"config": "C4_estimators", "estimator": label, "walkers": W, "n_per_chain": nmc, "series_GB": nbytes/1e9,
                          "walk_ms": t["walk_ms"], "estim_ms": t["estim_ms"], "estim_GBps": nbytes/(t["estim_ms"]*1e-3)/1e9,
                          "hbm_frac_of_measured": nbytes/(t["estim_ms"]*1e-3)/1e9/HBM_GBS, "walk_steps_per_s": W*nmc/(t["walk_ms"]*1e-3),
                          "walk_store_GBps": nbytes/(t["walk_ms"]*1e-3)/1e9, "avg": float(avg[0]), "err": float(err[0])}), flush=True)
    # one long chain block, n = 2^27 (BASELINE "Nmc=1e8 per chain block" -> next power of two), host data through mcig_estimate
    rng = np.random.default_rng(1)
    n = 1 << 27
    x = rng.normal(size=n)
    for i in range(0, 3):
        t0 = time.perf_counter()
        avg, err = m.estimate(m.EstimatorType.MJBlocker, x)
        wall = time.perf_counter() - t0
    print(json.dumps({"config": "C4_single_chain_2^27", "estimator": "MJBlocker", "wall_ms_incl_1GiB_H2D": 1e3*wall, "avg": float(avg[0]), "err": float(err[0]),
                      "expected_err": float(1/np.sqrt(n))}), flush=True)
    for label, est, xs in (("FCBlocker", m.EstimatorType.FCBlocker, x[:n - 1]), ("Uncorrelated", m.EstimatorType.Uncorrelated, x)):
        for i in range(0, 2):
            t0 = time.perf_counter()
            avg, err = m.estimate(est, xs)
            wall = time.perf_counter() - t0
        print(json.dumps({"config": "C4_single_chain_2^27", "estimator": label, "n": int(len(xs)), "wall_ms_incl_1GiB_H2D": 1e3*wall, "avg": float(avg[0]),
                          "err": float(err[0]), "expected_err": float(1/np.sqrt(n))}), flush=True)


def c5_mixed(W=65536, nmc=100000):
    mci = m.MCI(3)
    mci.setRngMode(0)
    mci.setSeed(5649871)
    mci.setNWalkers(W)
    mci.addSamplingFunction(m.ThreeDimGaussianPDF())
    mci.addObservable(m.XND(3), 0, 1)
    mci.addObservable(m.XSquared(), 1, 5)
    mci.addObservable(m.XYZSquared(), 5, 2)
    mci.setMRT2Step(1.0)
    for auto in (False, True, True):  # the first automatic run compiles / loads the calibration and equilibration kernel variants
        mci.integrate(1000, False, False)
        mci.setMRT2Step(1.0)
        t0 = time.perf_counter()
        avg, err = mci.integrate(nmc, auto, auto)
        wall = time.perf_counter() - t0
        t = mci.timings()
        print(json.dumps({"config": "C5_mixed", "auto_calibration_decorrelation": auto, "walkers": W, "nmc": nmc, "samples_per_s_total": W*nmc/(t["total_ms"]*1e-3),
                          "walk_ms": t["walk_ms"], "estim_ms": t["estim_ms"], "total_ms": t["total_ms"], "wall_ms": 1e3*wall, "launches": t["launches"], "find_ms": t["find_ms"], "decorr_ms": t["decorr_ms"], "jit_ms": t["jit_ms"],
                          "calibration_iterations": mci.getCalibrationIterations(), "step": mci.getMRT2Step(0), "acceptance": mci.getAcceptanceRate(), "avg": [float(v) for v in avg], "err": [float(v) for v in err]}), flush=True)


if __name__ == "__main__":
    which = sys.argv[1:] or ["c3", "c4", "c5"]
    if "c3" in which:
        c3_ndim("vec")
        c3_ndim("all")
        c3_ndim("multistep", ndims=(2, 4, 8, 16, 32))
    if "c3ms" in which:
        c3_ndim("multistep", ndims=(2, 4, 8, 16, 32, 64))
    if "c3all" in which:
        c3_ndim("all")
        c3_ndim("all", ndims=(128, 256, 1024), W=65536, nmc=1000)
        c3_ndim("multistep", ndims=(2, 4, 8, 16, 32, 64))
    if "c3vec" in which:
        c3_ndim("vec")
        c3_ndim("vec", ndims=(128, 256, 512, 1024), W=65536, nmc=4000)
    if "c3big" in which:  # the upper half of the reference's dimension sweeps (benchmark/bench_throughput_ndim_*: up to 1024)
        c3_ndim("vec", ndims=(128, 256, 512, 1024), W=16384, nmc=20000)
        c3_ndim("all", ndims=(128, 256, 512, 1024), W=16384, nmc=2000)
        c3_ndim("multistep", ndims=(64, 128, 256), W=16384, nmc=200)
    if "c3lanes" in which:  # lane-split walkers (placement 3) against the shared-memory placement
        for pl in (1, 3):
            c3_ndim("all", ndims=(32, 48, 64, 96, 128, 256), W=65536, nmc=2000, placement=pl)
    if "c4" in which:
        c4_estimators()
    if "c5" in which:
        c5_mixed()
