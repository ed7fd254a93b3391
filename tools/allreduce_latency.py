"""Latency of the path's only collective (a few doubles through parallel.allreduce_sum) under torchrun."""
import os
import sys
import time

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import numpy as np  # noqa: E402
import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

from mcintegratorplusplus_b200 import parallel  # noqa: E402

rank = parallel.MPIMCI.init()
torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
buf = np.ones(14)
for _ in range(20):
    parallel.allreduce_sum(buf)
dist.barrier()
torch.cuda.synchronize()
t0 = time.perf_counter()
for _ in range(200):
    parallel.allreduce_sum(buf)
t1 = time.perf_counter()
t = torch.ones(14, dtype=torch.float64, device="cuda")
torch.cuda.synchronize()
t2 = time.perf_counter()
for _ in range(200):
    dist.all_reduce(t)
torch.cuda.synchronize()
t3 = time.perf_counter()
if rank == 0:
    print("allreduce_sum (host in/out): %.1f us per call; raw dist.all_reduce (device tensor, async): %.1f us" % (1e6*(t1 - t0)/200, 1e6*(t3 - t2)/200))
parallel.MPIMCI.finalize()
