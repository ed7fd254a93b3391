"""CPU, world_size 2, gloo: the N>1 host logic — walker sharding, the cross-rank sum and the MPIMCI combination — with the
per-walker results supplied by the C oracle (one emulated MPI rank per walker, src/MPIMCI.cpp:83-92)."""
import ctypes as C
import os
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, seeds, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    sys.path.insert(0, os.path.join(ROOT, "tests"))
    import torch.distributed as dist
    import configs
    import orc
    from mcintegratorplusplus_b200 import parallel
    dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%d" % port, rank=rank, world_size=world)
    n, off = parallel.shard(len(seeds), rank, world)
    eng = orc.oracle()
    cfg = configs.make("full_mj")
    sums = np.zeros(8)
    for s in seeds[off:off + n]:  # this rank's walkers
        cfg.seed = s
        r = eng.run(cfg)
        sums[:4] += r["avg"]
        sums[4:] += np.array(r["err"])**2
    parallel.allreduce_sum(sums)
    avg, err = parallel.combine(sums, 4, len(seeds))
    q.put((rank, n, off, avg.tolist(), err.tolist()))
    dist.destroy_process_group()


def test_shard_partition():
    sys.path.insert(0, ROOT)
    from mcintegratorplusplus_b200 import parallel
    for total, world in ((65536, 8), (10, 4), (7, 8), (1, 1)):
        parts = [parallel.shard(total, r, world) for r in range(world)]
        assert sum(n for n, _ in parts) == total
        off = 0
        for n, o in parts:
            assert o == off
            off += n
        assert max(n for n, _ in parts) - min(n for n, _ in parts) <= 1


def test_two_rank_combination_matches_oracle(oracle):
    import torch.multiprocessing as mp
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import configs
    import orc
    seeds = [11, 22, 33, 44, 55]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    procs = [ctx.Process(target=_worker, args=(r, 2, port, seeds, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    res.sort()
    assert [(r[1], r[2]) for r in res] == [(3, 0), (2, 3)]
    # reference combination of the same 5 "ranks"
    cfg = configs.make("full_mj")
    f = oracle.lib.mcio_run_ranks
    f.restype = C.c_int
    f.argtypes = [C.POINTER(orc.Config), C.POINTER(C.c_uint64), C.c_int, C.POINTER(orc.Result), C.POINTER(orc.Result), C.POINTER(orc.Trace)]
    comb = orc.Result()
    assert f(C.byref(cfg), (C.c_uint64*len(seeds))(*seeds), len(seeds), C.byref(comb), None, None) == 0
    for rank, n, off, avg, err in res:
        assert np.allclose(avg, comb.avg[:4], rtol=1e-14, atol=0)
        assert np.allclose(err, comb.err[:4], rtol=1e-14, atol=0)
    assert res[0][3] == res[1][3]  # both ranks hold the same reduced result, like MPI_Allreduce
