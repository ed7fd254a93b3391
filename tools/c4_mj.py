"""C4 (Full accumulator + MJBlocker / FCBlocker / uncorrelated, 65536 chains x 2^14 samples) for a per-kernel launch list under ncu."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mcintegratorplusplus_b200 as m  # noqa: E402

for est in (m.EstimatorType.MJBlocker, m.EstimatorType.FCBlocker, m.EstimatorType.Uncorrelated):
    mci = m.MCI(3)
    mci.setRngMode(0)
    mci.setSeed(1337)
    mci.setNWalkers(65536)
    mci.addSamplingFunction(m.ThreeDimGaussianPDF())
    mci.addObservable(m.XSquared(), 1, 1, False, est)
    mci.setMRT2Step(1.0)
    for _ in range(2):
        mci.integrate(1 << 14, False, False)
    print(est, mci.timings())
