"""GPU, BASELINE.json full sizes (65536 walkers x 1e5 steps = 6.5e9 Metropolis steps per run, ~30 ms each): properties that do not
need an oracle run of that size — determinism, independence of the launch schedule, consistency of the walker combination,
the exact expectation, and exact scale covariance of the streaming estimators on a multi-GB series."""
import numpy as np
import pytest

import orc
from prod import build_mci

pytestmark = pytest.mark.gpu

W, NMC = 65536, 100000
SPEC = dict(ndim=3, seed=1337, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 0, 1)], nmc=NMC, steps=(1.0,))


def _run(mcig, dyn, blocksize=0):
    mci = build_mci(mcig, SPEC, nwalkers=W, mode=0)
    mci.setDynamicScheduling(dyn)
    if blocksize:
        mci.setBlockSize(blocksize)
    avg, err = mci.integrate(NMC, False, False)
    wavg, _ = mci.walkerResults()
    return avg, err, wavg[0].copy(), mci.getAcceptanceRate(), mci.crossWalkerError()[0], mci.sums(), np.array([mci.getX(walker=w) for w in (0, 12345, W - 1)])


def test_c2_full_size_schedule_independence_and_expectation(mcig):
    a = _run(mcig, dyn=1)            # persistent work-stealing kernel
    b = _run(mcig, dyn=0)            # static launch, 512-thread blocks
    c = _run(mcig, dyn=0, blocksize=128)
    d = _run(mcig, dyn=1)            # repeat: determinism
    for other in (b, c, d):
        assert np.array_equal(a[2], other[2]), "per-walker averages depend on the launch schedule"
        assert a[3] == other[3] and np.array_equal(a[6], other[6])
        assert a[0][0] == other[0][0]
    avg, err, wavg, acc, cw, sums, _ = a
    assert err[0] == 0.0                                   # Simple accumulator + Noop estimator (reference semantics)
    assert abs(wavg.sum() - sums[0]) <= 1e-9*abs(sums[0])    # combination = sum over walkers / W (src/MPIMCI.cpp:85-92)
    assert avg[0] == pytest.approx(sums[0]/W, rel=1e-15)
    assert abs(avg[0] - 0.5) < 4*cw and cw < 5e-5          # exact expectation 1/2 (test/ut2/main.cpp:65-91), 6.5e9 samples
    assert abs(acc - 0.505) < 2e-3                         # step 1.0 gives ~0.505 (benchmark/bench_integrate_mixed/main.cpp:44)
    # walkers are independent and identically distributed: their averages scatter like cw*sqrt(W)
    assert 0.8 < wavg.std(ddof=1)/(cw*np.sqrt(W)) < 1.2


def test_estimators_scale_covariance_on_large_series(mcig):
    """avg(a x + b) = a avg(x) + b and err(a x + b) = |a| err(x), exactly for a = 4 (power of two) and b = 0, to rounding for b != 0.
    Series: 4096 chains x 2^15 samples of the walk itself, pulled through obsData for a few walkers and re-estimated on the device."""
    spec = dict(ndim=3, seed=7, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 1, 1, False, orc.EST_MJBLOCKER)], nmc=1 << 15, steps=(1.0,))
    mci = build_mci(mcig, spec, nwalkers=4096, mode=0)
    avg, err = mci.integrate(1 << 15, False, False)
    wavg, werr = mci.walkerResults()
    for w in (0, 4095):
        x = mci.obsData(0, walker=w, nobs=1)[:, 0]
        for et in (orc.EST_MJBLOCKER, orc.EST_UNCORRELATED):
            a1, e1 = mcig.estimate(et, x)
            a4, e4 = mcig.estimate(et, 4.0*x)
            assert a4[0] == 4.0*a1[0] and e4[0] == 4.0*e1[0]
            ab, eb = mcig.estimate(et, x + 1.0)
            assert ab[0] == pytest.approx(a1[0] + 1.0, rel=1e-14) and eb[0] == pytest.approx(e1[0], rel=1e-9)
        a1, e1 = mcig.estimate(orc.EST_MJBLOCKER, x)
        assert a1[0] == pytest.approx(wavg[0, w], rel=1e-13) and e1[0] == pytest.approx(werr[0, w], rel=1e-9)
    # FCBlocker on a non-power-of-two prefix: scale covariance (event-driven kernel, n > 4096)
    x = mci.obsData(0, walker=1, nobs=1)[:30000, 0]
    a1, e1 = mcig.estimate(orc.EST_FCBLOCKER, x)
    a4, e4 = mcig.estimate(orc.EST_FCBLOCKER, 4.0*x)
    assert a4[0] == 4.0*a1[0] and e4[0] == 4.0*e1[0]


def test_full_accumulator_round_trip_at_scale(mcig):
    """Stored series = what the accumulator saw: per walker, mean of the Full series == Simple average of the same run (same
    streams), and block means of Block(16) == means of 16 consecutive Full samples."""
    base = dict(ndim=3, seed=99, pdf_id=orc.PDF_GAUSS3D, nmc=1 << 14, steps=(1.0,))
    m1 = build_mci(mcig, dict(base, obs=[(orc.OBS_XSQUARED, 0, 1), (orc.OBS_XSQUARED, 1, 1, False, orc.EST_UNCORRELATED),
                                         (orc.OBS_XSQUARED, 16, 1, False, orc.EST_UNCORRELATED)]), nwalkers=8192, mode=0)
    m1.setKeepSamples(True)
    avg, err = m1.integrate(1 << 14, False, False)
    wavg, _ = m1.walkerResults()
    assert np.allclose(wavg[0], wavg[1], rtol=1e-13, atol=0) and np.allclose(wavg[0], wavg[2], rtol=1e-13, atol=0)
    for w in (0, 8191):
        full = m1.obsData(1, walker=w, nobs=1)[:, 0]
        blocks = m1.obsData(2, walker=w, nobs=1)[:, 0]
        assert full.shape == (1 << 14,) and blocks.shape == (1 << 10,)
        ref_blocks = np.array([np.sum(full[16*i:16*(i + 1)]) for i in range(1 << 10)])*(1./16)
        assert np.allclose(blocks, ref_blocks, rtol=1e-15, atol=0)


def test_c4_full_size_chain_blocks_through_the_walk(mcig, oracle):
    """BASELINE configs[3] at its stated size: FullAccumulator + MJBlocker with n = 2^27 stored samples per chain block
    (benchmark/bench_estimators/main.cpp:31-50 rounds 1e8 up to the power of two src/MJBlocker.cpp:28-30 demands), staged in HBM by
    the walk kernel for as many chains as fit (128 chains = 137 GB; fewer on a box with less free memory, never below 16).
    One chain block is read back (1 GiB) and re-estimated by the oracle's own MJBlocker: mean 1e-12, error 1e-9."""
    import torch
    k = 27
    free, _ = torch.cuda.mem_get_info()
    W = 128
    while W > 16 and W*(8 << k)*1.06 + (2 << 30) > free:
        W //= 2
    if W*(8 << k)*1.06 + (2 << 30) > free:
        pytest.skip("not enough free HBM for 16 chain blocks of 2^27 samples")
    spec = dict(ndim=3, seed=2028, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 1, 1, False, orc.EST_MJBLOCKER)], nmc=1 << k, steps=(1.0,))
    mci = build_mci(mcig, spec, nwalkers=W, mode=0)
    avg, err = mci.integrate(1 << k, False, False)
    wavg, werr = mci.walkerResults()
    w = W - 3
    x = mci.obsData(0, walker=w, nobs=1)[:, 0]
    assert x.shape == (1 << k,)
    a, e = oracle.estimate(orc.EST_MJBLOCKER, x)
    assert wavg[0, w] == pytest.approx(a[0], rel=1e-12) and werr[0, w] == pytest.approx(e[0], rel=1e-9)
    assert abs(avg[0] - 0.5) < 4*err[0] and err[0] < 2e-5*np.sqrt(128/W)
    del mci


def test_c4_series_beyond_hbm_chunked_staging(mcig, oracle):
    """65536 chains x 2^20 stored samples = 550 GB of series on a 180 GB device (BASELINE configs[3] with device-side HBM staging): sampled,
    staged and folded chunk by chunk. One chain is re-run alone (Philox streams are keyed by the global walker id, so walker w of the big job
    and a one-walker job at global offset w are the same chain), its resident series goes through the oracle's MJBlocker
    (src/MJBlocker.cpp:127-154): mean 1e-12, error 1e-9."""
    W, k, w = 65536, 20, 40961
    spec = dict(ndim=3, seed=2029, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 1, 1, False, orc.EST_MJBLOCKER)], nmc=1 << k, steps=(1.0,))
    mci = build_mci(mcig, spec, nwalkers=W, mode=0)
    avg, err = mci.integrate(1 << k, False, False)
    assert mci.getStagingChunks() >= 4
    wavg, werr = mci.walkerResults()
    t = mci.timings()
    del mci
    one = build_mci(mcig, spec, mode=0)
    one.setNWalkers(1, global_offset=w, total=W)
    one.setKeepSamples(True)
    one.integrate(1 << k, False, False)
    x = one.obsData(0, walker=0, nobs=1)[:, 0]
    a, e = oracle.estimate(orc.EST_MJBLOCKER, x)
    assert wavg[0, w] == pytest.approx(a[0], rel=1e-12) and werr[0, w] == pytest.approx(e[0], rel=1e-9)
    assert abs(avg[0] - 0.5) < 4*err[0] and err[0] < 1e-5
    assert avg[0] == pytest.approx(wavg[0].sum()/W, rel=1e-13)
    assert t["total_ms"] < 5000
