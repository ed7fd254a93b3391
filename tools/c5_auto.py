"""C5 (mixed observables, automatic calibration + decorrelation) once warm, for a per-kernel launch list under ncu."""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import mcintegratorplusplus_b200 as m  # noqa: E402

W = int(sys.argv[1]) if len(sys.argv) > 1 else 65536
nmc = int(sys.argv[2]) if len(sys.argv) > 2 else 100000
mci = m.MCI(3)
mci.setRngMode(0)
mci.setSeed(5649871)
mci.setNWalkers(W)
mci.addSamplingFunction(m.ThreeDimGaussianPDF())
mci.addObservable(m.XND(3), 0, 1)
mci.addObservable(m.XSquared(), 1, 5)
mci.addObservable(m.XYZSquared(), 5, 2)
for rep in range(2):
    mci.setMRT2Step(1.0)
    avg, err = mci.integrate(nmc, True, True)
    print(mci.timings())
