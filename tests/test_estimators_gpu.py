"""GPU: device estimators (K3 MJBlocker, K4 uncorrelated, K5 FCBlocker, Noop) through mcig_estimate against the C oracle and
the reference's golden values, on the reference's own TestWalk data (test/ut1, benchmark/bench_estimators)."""
import numpy as np
import pytest

import configs
import orc
from conftest import fromhex

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("wname", sorted(configs.WALKS))
def test_estimators_match_reference(wname, mcig, oracle, golden_est):
    pdf, nmc, ndim, step, cp, seed = configs.WALKS[wname]
    datax, _, _, _, _ = oracle.testwalk(pdf, nmc, ndim, step, cp, seed)
    x = datax if ndim > 1 else datax[:, 0]
    g = golden_est[wname]
    for ename, et in (("uncorrelated", orc.EST_UNCORRELATED), ("correlated", orc.EST_CORRELATED), ("fcblocker", orc.EST_FCBLOCKER),
                      ("mjblocker", orc.EST_MJBLOCKER), ("noop", orc.EST_NOOP)):
        if ename not in g:
            continue
        avg, err = mcig.estimate(et, x)
        ga, ge = np.array(fromhex(g[ename]["avg"])), np.array(fromhex(g[ename]["err"]))
        assert np.allclose(avg, ga, rtol=1e-12, atol=1e-15), (ename, avg, ga)
        assert np.allclose(err, ge, rtol=1e-9, atol=1e-18), (ename, err, ge)


@pytest.mark.parametrize("wname", sorted(configs.WALKS))
def test_fixed_block_estimator_matches_reference(wname, mcig, oracle, golden_est):
    """One/MultiDimBlockEstimator (src/Estimators.cpp:59-80, 158-185) with the block count of the goldens (n/16 blocks)."""
    pdf, nmc, ndim, step, cp, seed = configs.WALKS[wname]
    datax, _, _, _, _ = oracle.testwalk(pdf, nmc, ndim, step, cp, seed)
    x = datax if ndim > 1 else datax[:, 0]
    g = golden_est[wname]["block16"]
    avg, err = mcig.estimate_blocks(x, nmc//16)
    assert np.allclose(avg, fromhex(g["avg"]), rtol=1e-12, atol=1e-15) and np.allclose(err, fromhex(g["err"]), rtol=1e-9, atol=1e-18)
    from mcintegratorplusplus_b200._capi import McigError
    with pytest.raises(McigError, match="n must be >= nblocks"):
        mcig.estimate_blocks(x[:10], 11)


def test_mjblocker_large_power_of_two_segments(mcig, oracle):
    """A series long enough to be time-split into segments (the HBM-streaming path) against the oracle's 13-pass algorithm."""
    rng = np.random.default_rng(7)
    n = 1 << 20
    e = rng.normal(size=(n, 2))
    x = np.empty_like(e)
    x[0] = e[0]
    for i in range(1, n):  # AR(1) correlated series
        x[i] = 0.9*x[i - 1] + e[i]
    avg_o, err_o = oracle.estimate(orc.EST_MJBLOCKER, x)
    avg, err = mcig.estimate(orc.EST_MJBLOCKER, x)
    assert np.allclose(avg, avg_o, rtol=1e-12, atol=1e-15)
    assert np.allclose(err, err_o, rtol=1e-9)
    avg_u, err_u = mcig.estimate(orc.EST_UNCORRELATED, x)
    avg_uo, err_uo = oracle.estimate(orc.EST_UNCORRELATED, x)
    assert np.allclose(avg_u, avg_uo, rtol=1e-12, atol=1e-15) and np.allclose(err_u, err_uo, rtol=1e-9)
    assert np.all(err > 2*err_u)  # blocking must see the correlation


def test_fcblocker_long_chains_time_split(mcig, oracle):
    """Few long chains whose length is not a power of two (the default estimator then is the FCBlocker, src/Estimators.cpp:278-286):
    the chains are split in time, block sums come from prefix differences across segments. Against the oracle's direct block sums."""
    rng = np.random.default_rng(11)
    n = 3*(1 << 19) + 17
    e = rng.normal(size=(n, 2))
    x = np.empty_like(e)
    x[0] = e[0]
    for i in range(1, n):
        x[i] = 0.8*x[i - 1] + e[i] + 0.25
    for data in (x, x[:, 0].copy(), x[:100003]):
        avg_o, err_o = oracle.estimate(orc.EST_FCBLOCKER, data)
        avg, err = mcig.estimate(orc.EST_FCBLOCKER, data)
        assert np.allclose(avg, avg_o, rtol=1e-12, atol=1e-15), (avg, avg_o)
        assert np.allclose(err, err_o, rtol=1e-9), (err, err_o)
        avg_c, err_c = mcig.estimate(orc.EST_CORRELATED, data)  # dispatches to the FCBlocker: n is not a power of two
        assert np.array_equal(avg_c, avg) and np.array_equal(err_c, err)


@pytest.mark.parametrize("n", [50, 51, 99, 100, 127, 198, 199, 346, 384, 385, 1000, 4096])
def test_fcblocker_short_series_exact_paths(n, mcig, oracle):
    """Short series (the 100- and 346-sample chunks of the decorrelation loop, src/MCIntegrator.cpp:193) run through the exact kernels: one pass per
    distinct block length shared by a group of partitions, partition groups spread over the warps of a block for n <= 198, one thread per chain
    up to 4096. Every block sum keeps the reference's order, so the oracle's values are reproduced to the last bits; 70 chains cross the
    32-chain tiles of the shared-memory kernels."""
    rng = np.random.default_rng(100 + n)
    e = rng.normal(size=(n, 70))
    x = np.empty_like(e)
    x[0] = e[0]
    for i in range(1, n):
        x[i] = 0.7*x[i - 1] + e[i] + 0.5
    avg_o, err_o = oracle.estimate(orc.EST_FCBLOCKER, x)
    avg, err = mcig.estimate(orc.EST_FCBLOCKER, x)
    assert np.allclose(avg, avg_o, rtol=1e-14, atol=0.), np.max(np.abs(avg - avg_o))
    assert np.allclose(err, err_o, rtol=1e-13, atol=0.), np.max(np.abs(err/err_o - 1))
    one_o = oracle.estimate(orc.EST_FCBLOCKER, x[:, 3].copy())
    one = mcig.estimate(orc.EST_FCBLOCKER, x[:, 3].copy())  # the one-dimensional estimator divides where the multi-dimensional one multiplies (src/Estimators.cpp:69 vs :175)
    assert np.allclose(one[0], one_o[0], rtol=1e-14, atol=0.) and np.allclose(one[1], one_o[1], rtol=1e-13, atol=0.)


def test_constant_series_defined_error(mcig, oracle):
    """Constval-like data (test/main.cpp:82-88). An exactly representable constant has zero variance at every level: the
    reference then reads out of bounds; here and in the oracle err is defined as 0. A constant
    like 1.3 leaves rounding residue in x - mean and behaves like ordinary data: compare with the oracle."""
    avg, err = mcig.estimate(orc.EST_MJBLOCKER, np.full(1024, 1.5))
    assert avg[0] == 1.5 and err[0] == 0.0
    avg, err = mcig.estimate(orc.EST_MJBLOCKER, np.full(1024, 1.3))
    avg_o, err_o = oracle.estimate(orc.EST_MJBLOCKER, np.full(1024, 1.3))
    assert avg[0] == avg_o[0] and err[0] == pytest.approx(err_o[0], rel=1e-9, abs=1e-18)


def test_estimator_argument_errors(mcig):
    from mcintegratorplusplus_b200._capi import McigError
    with pytest.raises(McigError, match="power of two"):
        mcig.estimate(orc.EST_MJBLOCKER, np.arange(100.))
    with pytest.raises(McigError, match=">= 50"):
        mcig.estimate(orc.EST_FCBLOCKER, np.arange(40.))
    with pytest.raises(McigError, match="larger than 1"):
        mcig.estimate(orc.EST_UNCORRELATED, np.arange(1.))


def test_c4_full_accumulator_mjblocker_through_the_walk(mcig, oracle):
    """BASELINE configs[3] shape through the walk itself (not host data): FullAccumulator + MJBlocker on the 3-D Gaussian / x^2 integrand,
    128 chains x 2^22 stored samples (4.3 GB staged in HBM by the walk kernel). Sampled chains are pulled back and re-estimated by the
    oracle's 13-pass MJBlocker (src/MJBlocker.cpp:127-154): mean to 1e-12, error to 1e-9; the combination is src/MPIMCI.cpp:85-92."""
    W, k = 128, 22
    spec = dict(ndim=3, seed=2027, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 1, 1, False, orc.EST_MJBLOCKER)], nmc=1 << k, steps=(1.0,))
    from prod import build_mci
    mci = build_mci(mcig, spec, nwalkers=W, mode=0)
    avg, err = mci.integrate(1 << k, False, False)
    wavg, werr = mci.walkerResults()
    for w in (0, 77, W - 1):
        x = mci.obsData(0, walker=w, nobs=1)[:, 0]
        assert x.shape == (1 << k,)
        a, e = oracle.estimate(orc.EST_MJBLOCKER, x)
        assert wavg[0, w] == pytest.approx(a[0], rel=1e-12)
        assert werr[0, w] == pytest.approx(e[0], rel=1e-9)
        assert e[0] > 0
    assert avg[0] == pytest.approx(wavg[0].sum()/W, rel=1e-13)
    assert err[0] == pytest.approx(np.sqrt((werr[0]**2).sum())/W, rel=1e-12)
    assert abs(avg[0] - 0.5) < 4*err[0]


def test_chunked_staging_equals_resident_series(mcig, oracle, monkeypatch):
    """A stored series that does not fit HBM is sampled, staged and folded chunk by chunk (include/mci/FullAccumulator.hpp:11-13 warns about
    the memory; src/MJBlocker.cpp:46-154 needs the global mean, which the chunks only know at the end). Forced here on a small run with
    MCIG_CHUNK_BYTES: per-walker means must be the resident run's to rounding, MJBlocker errors to 1e-9 (the folded level sums are centred
    on the first chunk's mean and corrected exactly), Noop / fused observables and the chain itself must not change at all; one chain is also
    checked against the oracle's own MJBlocker."""
    from prod import build_mci
    W, k = 2048, 16
    spec = dict(ndim=3, seed=77, pdf_id=orc.PDF_GAUSS3D, nmc=1 << k, steps=(1.0,),
                obs=[(orc.OBS_XSQUARED, 1, 1, False, orc.EST_MJBLOCKER), (orc.OBS_XYZSQUARED, 4, 2, False, orc.EST_MJBLOCKER), (orc.OBS_XND, 0, 1),
                     (orc.OBS_XYZSQUARED, 1, 4, False, orc.EST_UNCORRELATED), (orc.OBS_XSQUARED, 1, 8, False, orc.EST_NOOP)])
    out = []
    for budget in (None, 96 << 20):
        if budget:
            monkeypatch.setenv("MCIG_CHUNK_BYTES", str(budget))
        else:
            monkeypatch.delenv("MCIG_CHUNK_BYTES", raising=False)
        mci = build_mci(mcig, spec, nwalkers=W, mode=0)
        mci.setKeepSamples(not budget)
        mci.integrate(1 << 12, False, False)   # chains continue from wherever the first call left them
        avg, err = mci.integrate(1 << k, False, False)
        wavg, werr = mci.walkerResults()
        x = mci.obsData(0, walker=5, nobs=1)[:, 0] if not budget else None
        out.append((avg.copy(), err.copy(), wavg.copy(), werr.copy(), mci.getAcceptanceRate(), [list(mci.getX(walker=w)) for w in (0, 999, W - 1)], mci.getStagingChunks(), x))
        if budget:
            from mcintegratorplusplus_b200._capi import McigError
            with pytest.raises(McigError, match="were not stored"):
                mci.obsData(0, walker=5, nobs=1)
    a, b = out
    assert a[6] == 0 and b[6] >= 4
    assert a[4] == b[4] and a[5] == b[5]
    assert np.allclose(a[2], b[2], rtol=1e-13, atol=1e-16)
    assert np.allclose(a[3], b[3], rtol=1e-9, atol=1e-18)
    assert np.array_equal(a[2][4:7], b[2][4:7]) and np.array_equal(a[2][10], b[2][10])  # Simple and Noop legs: bit-identical
    assert np.allclose(a[0], b[0], rtol=1e-13, atol=1e-16) and np.allclose(a[1], b[1], rtol=1e-9, atol=1e-18)
    av, er = oracle.estimate(orc.EST_MJBLOCKER, a[7])
    assert b[2][0, 5] == pytest.approx(av[0], rel=1e-12) and b[3][0, 5] == pytest.approx(er[0], rel=1e-9)
    monkeypatch.setenv("MCIG_CHUNK_BYTES", str(96 << 20))
    mci = build_mci(mcig, dict(spec, obs=[(orc.OBS_XSQUARED, 1, 1)], nmc=60000*8), nwalkers=W, mode=0)  # Correlated on a non power of two: FCBlocker
    from mcintegratorplusplus_b200._capi import McigError
    with pytest.raises(McigError, match="chunked staging"):
        mci.integrate(60000*8, False, False)
