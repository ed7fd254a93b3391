"""GPU, needs >= 2 B200s on the box (skipped otherwise; run with `gpurun --gpus 2 -- python -m pytest tests/test_multi_gpu.py -m gpu`):
the collective INSIDE the library (include/mcig.h: mcig_comm_*, NCCL over NVLink, on the engine's stream) against the single-process job
and against the host-callback path, from Python (torchrun) and from a reference-style C++ SPMD program (include/mci/MPIMCI.hpp).
Reference: src/MPIMCI.cpp:27-93, src/MCIntegrator.cpp:21-34, 131-138, 203-230, examples/ex_mpi/main.cpp."""
import json
import os
import subprocess
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "mcintegratorplusplus_b200")


def _ngpu():
    import torch
    return torch.cuda.device_count()


def _check(single, sharded):
    assert sharded["step"] == single["step"], "sharding changed the calibrated step size"
    assert sharded["iters"] == single["iters"] and sharded["chunks"] == single["chunks"]
    assert sharded["acc"] == pytest.approx(single["acc"], abs=5e-3)  # getAcceptanceRate is per process (its own walkers), as per rank in the reference
    for k in ("avg", "avg2"):
        assert np.allclose(sharded[k], single[k], rtol=1e-11, atol=1e-13), (k, sharded[k], single[k])
    for k in ("err", "err2"):
        assert np.allclose(sharded[k], single[k], rtol=1e-9, atol=1e-15), (k, sharded[k], single[k])


@pytest.mark.parametrize("world", [2, 4, 8])
def test_sharded_job_equals_single_process_job(world, mcig):
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", str(world), "--master-addr", "127.0.0.1", "--master-port", str(29600 + world),
           os.path.join(ROOT, "tests", "multi_gpu_check.py")]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    line = [l for l in r.stdout.splitlines() if l.startswith("RESULT ")][-1]
    out = json.loads(line[len("RESULT "):])
    assert out["ranks_agree"]
    _check(out["single"], out["nccl"])      # device-resident loops + ncclAllReduce on the stream
    _check(out["single"], out["callback"])  # host loops + host all-reduce callback
    assert 0.4 < out["single"]["acc"] < 0.6 and out["single"]["iters"] >= 3 and out["single"]["chunks"] >= 2
    assert all(abs(a - 0.5) < 5*e for a, e in zip(out["single"]["avg"][3:], out["single"]["err"][3:]))


@pytest.mark.parametrize("world", [2, 8])
def test_cpp_mpimci_program_over_several_gpus(world, mcig, tmp_path):
    """A reference-style main() (MPIMCI::init / setSeed / integrate / finalize) started once per GPU by tools/mcirun.sh."""
    if _ngpu() < world:
        pytest.skip("needs %d GPUs" % world)
    exe = str(tmp_path / "test_mpimci")
    subprocess.run(["g++", "-std=c++14", "-O1", "-Wall", "-Wextra", "-I" + os.path.join(ROOT, "include"), os.path.join(ROOT, "tests", "cpp", "test_mpimci.cpp"),
                    "-L" + PKG, "-lmcig", "-Wl,-rpath," + PKG, "-o", exe], check=True, capture_output=True, text=True)
    seeds = tmp_path / "rseed.txt"
    seeds.write_text(" ".join(str(1000003*i + 17) for i in range(9000)))
    res = {}
    for n in (1, world):
        env = dict(os.environ, MASTER_PORT=str(29700 + n))
        r = subprocess.run([os.path.join(ROOT, "tools", "mcirun.sh"), str(n), exe, "8192", "20000", str(seeds)], capture_output=True, text=True, timeout=900, env=env)
        assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
        first = [l for l in r.stdout.splitlines() if l.startswith("MPIMCI ranks")][0].split()
        second = [l for l in r.stdout.splitlines() if l.startswith("MPIMCI second")][0].split()
        assert int(first[2]) == n
        res[n] = (float(first[4]), float(first[6]), np.array(first[7:], dtype=float), np.array(second[2:], dtype=float))
    assert res[1][0] == res[world][0], "sharding changed the calibrated step size"
    assert res[1][1] == pytest.approx(res[world][1], abs=5e-3)
    assert np.allclose(res[1][2], res[world][2], rtol=1e-9, atol=1e-13) and np.allclose(res[1][3], res[world][3], rtol=1e-9, atol=1e-13)
    avg = res[world][2][0::2]
    err = res[world][2][1::2]
    assert all(abs(a - 0.5) < 5*e for a, e in zip(avg[3:], err[3:]))
