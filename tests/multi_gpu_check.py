"""One rank of the N-GPU consistency check (launched by tests/test_multi_gpu.py through torchrun): the same job — mixed observables with
automatic step calibration and decorrelation (BASELINE configs[4]) — run (a) sharded over the ranks with the library's own NCCL collective
(device-resident control loops), (b) sharded with the host all-reduce callback (host loops), (c) by rank 0 alone over all walkers.
Sharding must not change the control decisions: calibrated step, number of calibration iterations and decorrelation chunks are identical;
the estimates agree to rounding (the cross-rank sum has another order than the single-process tree)."""
import json
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import mcintegratorplusplus_b200 as m  # noqa: E402
from mcintegratorplusplus_b200 import parallel  # noqa: E402

TOTAL, NMC = 8192, 20000


def make(local, n, off, total):
    mci = m.MCI(3, device=local)
    mci.setRngMode(0)
    mci.setSeed(5649871)
    mci.setNWalkers(n, global_offset=off, total=total)
    mci.setX([1.0, -0.5, 0.25])
    mci.addSamplingFunction(m.ThreeDimGaussianPDF())
    mci.addObservable(m.XND(3), 0, 1)
    mci.addObservable(m.XSquared(), 1, 5)
    mci.addObservable(m.XYZSquared(), 5, 2)
    mci.setMRT2Step(0.2)
    return mci


def run(mci):
    avg, err = mci.integrate(NMC, True, True)
    avg2, err2 = mci.integrate(NMC, False, False)  # chains continue
    return {"avg": list(map(float, avg)), "err": list(map(float, err)), "avg2": list(map(float, avg2)), "err2": list(map(float, err2)), "step": mci.getMRT2Step(0),
            "acc": mci.getAcceptanceRate(), "iters": mci.getCalibrationIterations(), "chunks": mci.getDecorrelationChunks()}


def main():
    local = int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    rank, world = dist.get_rank(), dist.get_world_size()
    n, off = parallel.shard(TOTAL, rank, world)
    assert parallel.init_comm(local) == (rank, world)
    a = make(local, n, off, TOTAL)
    a.attachComm()
    ra = run(a)
    b = make(local, n, off, TOTAL)
    b.setAllreduce(lambda buf: parallel.allreduce_sum(buf))
    rb = run(b)
    out = {"nccl": ra, "callback": rb}
    if rank == 0:
        c = make(local, TOTAL, 0, TOTAL)
        out["single"] = run(c)
    # every rank must hold the same combined result
    t = torch.tensor(ra["avg"] + ra["err"] + [ra["step"]], dtype=torch.float64, device="cuda")
    lo, hi = t.clone(), t.clone()
    dist.all_reduce(lo, op=dist.ReduceOp.MIN)
    dist.all_reduce(hi, op=dist.ReduceOp.MAX)
    out["ranks_agree"] = bool(torch.equal(lo, hi))
    dist.barrier()
    parallel.finalize_comm()
    dist.destroy_process_group()
    if rank == 0:
        print("RESULT " + json.dumps(out))


if __name__ == "__main__":
    main()
