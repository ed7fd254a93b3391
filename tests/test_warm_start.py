"""GPU: warm start. A process that calls prebuild() must find every kernel of its configuration ready at its FIRST integrate(): the JIT-compiled walk
kernels come from the on-disk cubin cache (filled by an earlier process), and the estimator / control kernels are already on the device instead of
being loaded lazily at their first launch (round 1: 54 ms for the first estimator stage against 4.3 ms in steady state, 401 ms of NVRTC). warmup(nmc)
additionally rehearses the call (undone afterwards), so that its device buffers exist at their final size: the first estimator stage then takes what every later one takes."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

SCRIPT = r"""
import json, sys, time
sys.path.insert(0, %r)
import mcintegratorplusplus_b200 as m
mci = m.MCI(3)
mci.setRngMode(0); mci.setSeed(11); mci.setNWalkers(16384)
mci.addSamplingFunction(m.ThreeDimGaussianPDF())
mci.addObservable(m.XND(3), 0, 1)                                            # Simple
mci.addObservable(m.XSquared(), 1, 5, True, m.EstimatorType.Correlated)      # Full -> MJBlocker / FCBlocker
mci.addObservable(m.XYZSquared(), 5, 2, True, m.EstimatorType.Uncorrelated)  # Block
mci.addObservable(m.XSquared(), 1, 1, False, m.EstimatorType.FCBlocker)
t0 = time.perf_counter()
mci.warmup(20000, True, True)  # prebuild() + an undone rehearsal of the call below
prebuild_s = time.perf_counter() - t0
x0, g0, s0 = list(mci.getX()), mci.getStreamPosition(), mci.getMRT2Step(0)
out = []
for _ in range(4):
    avg, err = mci.integrate(20000, True, True)
    t = mci.timings()
    t["avg3"] = float(avg[3])
    out.append(t)
print(json.dumps({"prebuild_s": prebuild_s, "calls": out, "state": [x0, g0, s0]}))
""" % ROOT


def _run(cache):
    env = dict(os.environ, MCIG_CACHE_DIR=cache)
    r = subprocess.run([sys.executable, "-c", SCRIPT], capture_output=True, text=True, timeout=600, env=env)
    assert r.returncode == 0, r.stdout + r.stderr
    return json.loads(r.stdout.strip().splitlines()[-1])


def test_second_process_starts_warm(tmp_path):
    cache = str(tmp_path / "cubins")
    first = _run(cache)   # fills the cache (NVRTC runs here, inside prebuild)
    files = {f: os.path.getmtime(os.path.join(cache, f)) for f in os.listdir(cache)}
    assert len(files) >= 3  # main, calibration and equilibration variants of the walk kernel
    second = _run(cache)
    # the second process compiled nothing: same files, untouched (its prebuild time is CUDA context creation + module loads, which varies by seconds
    # between processes on the same box and is not asserted)
    assert {f: os.path.getmtime(os.path.join(cache, f)) for f in os.listdir(cache)} == files
    assert [c["avg3"] for c in first["calls"]] == [c["avg3"] for c in second["calls"]]  # same streams in both processes, rehearsed or not
    for proc in (first, second):  # prebuild() did the compiling and the loading: no integrate call of either process pays for it
        calls = proc["calls"]
        assert all(c["jit_ms"] == 0 for c in calls), [c["jit_ms"] for c in calls]
        assert proc["state"] == [[0.0, 0.0, 0.0], 0, 0.05]  # the rehearsal left no trace: start position, Philox cursor, step size
        steady = min(c["estim_ms"] for c in calls[1:])
        assert calls[0]["estim_ms"] <= 1.5*steady + 0.3, (calls[0]["estim_ms"], steady)
