"""CPU, only where oracle/_ref/libmci_ref.so exists (built from /root/reference by `make -C oracle ref`): the C oracle
against the live reference, bit for bit, including full accept sequences and draw streams, plus multi-rank emulation."""
import ctypes as C

import numpy as np
import pytest

import configs
import orc

pytestmark = pytest.mark.skipif(not orc.have_ref(), reason="oracle/_ref not built (needs /root/reference)")


@pytest.fixture(scope="module")
def ref():
    return orc.ref()


@pytest.mark.parametrize("name", sorted(n for n in configs.RUNS if configs.in_oracle(n)))
def test_live_reference_bit_exact(name, oracle, ref):
    cfg = configs.make(name)
    a = ref.run(cfg, trace=True)
    b = oracle.run(cfg, trace=True)
    for f in ("avg", "err", "acc_rate", "x_final", "steps_final", "n_acc", "n_rej"):
        assert a[f] == b[f], f
    assert np.array_equal(a["accepted"], b["accepted"])
    if a["n_draws"] >= 0:  # Gaussian draws are not exported by the harness
        assert a["n_draws"] == b["n_draws"] and np.array_equal(a["draws"], b["draws"])


def test_export_draws_matches_trace(oracle):
    cfg = configs.make("vec_exp4")
    r = oracle.run(cfg, trace=True)
    f = oracle.lib.mcio_export_draws
    f.restype = C.c_int64
    f.argtypes = [C.POINTER(orc.Config), C.c_uint64, C.c_int64, C.c_int64, C.POINTER(C.c_double), C.c_int64]
    out = np.zeros(r["n_draws"])
    n = f(C.byref(cfg), cfg.seed, 0, cfg.nmc, out.ctypes.data_as(C.POINTER(C.c_double)), len(out))
    assert n == r["n_draws"] and np.array_equal(out, r["draws"])


def test_ranks_combination(oracle, ref):
    """R emulated MPI ranks without in-loop collectives == R independent reference runs combined by src/MPIMCI.cpp:85-92."""
    cfg = configs.make("full_mj")
    seeds = [11, 22, 33, 44]
    f = oracle.lib.mcio_run_ranks
    f.restype = C.c_int
    f.argtypes = [C.POINTER(orc.Config), C.POINTER(C.c_uint64), C.c_int, C.POINTER(orc.Result), C.POINTER(orc.Result), C.POINTER(orc.Trace)]
    comb = orc.Result()
    per = (orc.Result*len(seeds))()
    sarr = (C.c_uint64*len(seeds))(*seeds)
    assert f(C.byref(cfg), sarr, len(seeds), C.byref(comb), per, None) == 0
    avgs, errs = [], []
    for i, s in enumerate(seeds):
        cfg.seed = s
        r = ref.run(cfg)
        assert r["avg"] == list(per[i].avg[:4]) and r["err"] == list(per[i].err[:4])
        avgs.append(r["avg"])
        errs.append(r["err"])
    avgs, errs = np.array(avgs), np.array(errs)
    for j in range(4):
        a = sum(avgs[:, j])/len(seeds)  # left-to-right like the emulation
        assert comb.avg[j] == pytest.approx(a, rel=1e-15)
        assert comb.err[j] == pytest.approx(np.sqrt(sum(errs[:, j]**2))/len(seeds), rel=1e-15)
