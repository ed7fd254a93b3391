"""Short run of the headline walk kernel for ncu (one launch is replayed ~40x under --set full: keep it small)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import mcintegratorplusplus_b200 as m  # noqa: E402

if os.environ.get("MCIG_TOOL_RK_IMM") == "1":  # experiment: Philox round keys of seed 1337 as immediates (MCIG_RK_IMM, see mcig_device.cuh)
    keys = []
    for r in range(10):
        keys += [(1337 + r*0x9E3779B9) & 0xffffffff, (r*0xBB67AE85) & 0xffffffff]
    os.environ["MCIG_JIT_DEFINES"] = ";".join(x for x in (os.environ.get("MCIG_JIT_DEFINES", ""), "MCIG_RK_IMM=" + ",".join("0x%xu" % k for k in keys)) if x)

nmc = int(sys.argv[1]) if len(sys.argv) > 1 else 4000
W = int(sys.argv[2]) if len(sys.argv) > 2 else 65536
bs = int(sys.argv[3]) if len(sys.argv) > 3 else 0
dyn = int(sys.argv[4]) if len(sys.argv) > 4 else -1
mci = m.MCI(3)
mci.setRngMode(0)
mci.setSeed(1337)
mci.setNWalkers(W)
mci.addSamplingFunction(m.ThreeDimGaussianPDF())
mci.addObservable(m.XSquared(), 0, 1)
mci.setMRT2Step(1.0)
if bs:
    mci.setBlockSize(bs)
mci.setDynamicScheduling(dyn)
for _ in range(3):
    avg, err = mci.integrate(nmc, False, False)
t = mci.timings()
print("dyn=%d " % dyn, end="")
print("W=%d nmc=%d bs=%d walk %.3f ms -> %.4e steps/s, avg %.6f acc %.4f" % (W, nmc, bs, t["walk_ms"], W*nmc/(t["walk_ms"]*1e-3), avg[0], mci.getAcceptanceRate()))
