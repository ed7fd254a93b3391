"""C5 of BASELINE.json (bench_integrate_mixed with automatic step calibration + decorrelation, walkers sharded over N GPUs, one NCCL
all-reduce of the accumulator sums per integrate plus the rate / estimate sums of the automatic routines) under torchrun:
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P tools/c5_multi.py
Rank 0 prints one JSON line. Timing: barrier + device synchronisation on both sides, max over ranks."""
import json
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import mcintegratorplusplus_b200 as m  # noqa: E402
from mcintegratorplusplus_b200 import parallel  # noqa: E402

W_PER_GPU = 65536
NMC = int(sys.argv[1]) if len(sys.argv) > 1 else 100000

rank = parallel.MPIMCI.init()
world = parallel.MPIMCI.size()
local = int(os.environ.get("LOCAL_RANK", 0))
torch.cuda.set_device(local)
mci = m.MCI(3, device=local)
mci.setRngMode(0)
mci.setSeed(5649871)
parallel.install(mci, W_PER_GPU*world)
mci.addSamplingFunction(m.ThreeDimGaussianPDF())
mci.addObservable(m.XND(3), 0, 1)          # benchmark/bench_integrate_mixed/main.cpp:31-41
mci.addObservable(m.XSquared(), 1, 5)
mci.addObservable(m.XYZSquared(), 5, 2)


def barrier():
    if world > 1:
        dist.barrier()
    torch.cuda.synchronize()


out = []
for rep in range(3):  # the first run compiles / loads the calibration and equilibration kernel variants
    mci.setMRT2Step(1.0)
    barrier()
    t0 = time.perf_counter()
    avg, err = mci.integrate(NMC, True, True)
    barrier()
    wall = time.perf_counter() - t0
    t = torch.tensor([wall], dtype=torch.float64, device="cuda")
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    out.append((float(t.item()), avg, err, mci.timings()))
wall, avg, err, tm = out[-1]
if rank == 0:
    exact = np.array([0., 0., 0., .5, .5, .5, .5])
    cw = np.asarray(mci.crossWalkerError())/np.sqrt(world)  # Simple accumulators report error 0 (Noop estimator): use the cross-walker spread
    pull = np.abs(avg - exact)/np.where(err > 0., err, cw)
    print(json.dumps({"config": "C5_mixed_auto", "n_gpus": world, "walkers_total": W_PER_GPU*world, "nmc": NMC, "wall_ms": 1e3*wall,
                      "samples_per_s": W_PER_GPU*world*NMC/wall, "walk_ms": tm["walk_ms"], "estim_ms": tm["estim_ms"], "find_ms": tm["find_ms"],
                      "decorr_ms": tm["decorr_ms"], "launches": tm["launches"], "total_ms": tm["total_ms"], "step": mci.getMRT2Step(0), "acceptance": mci.getAcceptanceRate(),
                      "avg": [float(v) for v in avg], "err": [float(v) for v in err], "max_pull_sigma": float(pull.max())}), flush=True)
    assert pull.max() < 5., pull
parallel.MPIMCI.finalize()
