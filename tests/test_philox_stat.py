"""GPU: production (Philox) mode. Integral estimates must agree with the reference's (oracle, pinned) within 3 combined
standard errors and with the exact expectations the reference's own tests use (test/ut2/main.cpp:65-91: <x^2> = 0.5)."""
import numpy as np
import pytest

import configs
import orc
from prod import build_mci

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("mode", [0, 1])
def test_c2_gaussian_philox_matches_reference_statistics(mode, mcig, oracle):
    ref = oracle.run(configs.make("auto_default"))  # reference estimate 0.4963 +- 0.0049
    spec = dict(ndim=3, seed=1337, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 0, 1)], nmc=8192, steps=(1.0,))
    mci = build_mci(mcig, spec, nwalkers=8192, mode=mode)
    mci.integrate(2000, False, False)  # warm-up like benchmark/bench_integrate_mixed/main.cpp:46
    avg, err = mci.integrate(8192, False, False)
    cw = mci.crossWalkerError()
    assert err[0] == 0.0  # Simple accumulator + Noop estimator: reference semantics
    assert 0.45 < mci.getAcceptanceRate() < 0.56
    assert abs(avg[0] - ref["avg"][0]) < 3*np.hypot(cw[0], ref["err"][0])
    assert abs(avg[0] - 0.5) < 4*cw[0]
    assert cw[0] < 2e-3


def test_philox_results_independent_of_sharding_and_launch_chunking(mcig):
    """Streams are keyed by (seed, GLOBAL walker id, step): a shard [128, 256) of 512 walkers reproduces walkers 128..255 bit for bit."""
    spec = dict(ndim=3, seed=99, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 0, 1)], nmc=1000, steps=(1.0,))
    full = build_mci(mcig, spec, nwalkers=512, mode=0)
    full.integrate(1000, False, False)
    wavg_full, _ = full.walkerResults()
    shard = mcig.MCI(3)
    shard.setRngMode(0)
    shard.setNWalkers(128, global_offset=128, total=512)
    shard.setSeed(99)
    shard.addSamplingFunction(mcig.ThreeDimGaussianPDF())
    shard.addObservable(mcig.XSquared(), 0, 1)
    shard.setMRT2Step(1.0)
    shard.integrate(1000, False, False)
    wavg_shard, _ = shard.walkerResults()
    assert np.array_equal(wavg_shard[0], wavg_full[0, 128:256])
    # two launches of 500 steps continue the same streams as one launch of 1000
    two = build_mci(mcig, spec, nwalkers=512, mode=0)
    two.integrate(500, False, False)
    two.integrate(500, False, False)
    assert np.array_equal(two.getX(walker=7), full.getX(walker=7))


@pytest.mark.parametrize("name", ["mixed", "vec_exp4", "ms_sub_ut5", "ndim_vec16", "exbasic_2", "nopdf_box", "ut4_fixed", "gauss_all_auto", "gauss_vec5"])
def test_families_within_three_sigma_of_reference(name, mcig, oracle):
    spec = dict(configs.RUNS[name])
    ref = oracle.run(configs.make(name))
    W = 256
    mci = build_mci(mcig, spec, nwalkers=W, mode=0)
    avg, err = mci.integrate(spec["nmc"], spec.get("do_find", False), spec.get("do_decorr", False))
    cw = mci.crossWalkerError()
    for j in range(len(avg)):
        e_ref = ref["err"][j]
        if e_ref == 0.0:  # Noop estimator in the reference: use the GPU's cross-walker error scaled to one chain
            e_ref = cw[j]*np.sqrt(W)
        assert abs(avg[j] - ref["avg"][j]) < 3.5*np.hypot(cw[j], e_ref) + 1e-12, (name, j, avg[j], ref["avg"][j], cw[j], e_ref)


def test_auto_calibration_reaches_target_rate(mcig):
    spec = dict(configs.RUNS["auto_default"])
    mci = build_mci(mcig, spec, nwalkers=4096, mode=0)
    avg, err = mci.integrate(4096, True, True)
    assert abs(mci.getAcceptanceRate() - 0.5) < 0.03
    assert 0.7 < mci.getMRT2Step(0) < 1.3  # reference calibrates to 0.987
    assert abs(avg[0] - 0.5) < 4*mci.crossWalkerError()[0]
    assert err[0] > 0


@pytest.mark.parametrize("name,exact", [("ms_sub16", 0.0), ("ms_nosub8", 0.0), ("ndim_all16", 0.0), ("vec_exp4", 0.0)])
def test_symmetric_expectations_exact(name, exact, mcig):
    """<x_i> = 0 for the symmetric fixtures pdfs (XND observable): a known answer independent of the reference's own
    (block-correlated, hence optimistic) error bars; checked with the cross-walker standard error of 512 chains."""
    spec = dict(configs.RUNS[name])
    spec["obs"] = [(orc.OBS_XND, 0, 1)]
    W = 512
    mci = build_mci(mcig, spec, nwalkers=W, mode=0)
    mci.integrate(2000, False, False)  # decorrelate from the common start
    avg, _ = mci.integrate(4000, False, False)
    cw = mci.crossWalkerError()
    assert np.all(np.abs(avg - exact) < 4.5*cw + 1e-12), (avg, cw)


PREFILTER_SPECS = {
    "all3_reg": dict(ndim=3, seed=2024, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 0, 1)], nmc=20000, steps=(1.0,), x0=(0.3, -0.1, 0.2)),
    "vec8_smem": dict(ndim=8, seed=2025, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_X2SUM, 0, 1)], nmc=20000, move_type=orc.MOVE_VEC, veclen=1, steps=(3.0,)),
    "vec6_v2_gauss_smem": dict(ndim=6, seed=2026, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2SUM, 0, 1)], nmc=20000, move_type=orc.MOVE_VEC, veclen=2, steps=(0.9,)),
    "all32_smem": dict(ndim=32, seed=2027, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2SUM, 0, 1)], nmc=4000, steps=(0.2,)),
}


@pytest.mark.parametrize("case", sorted(PREFILTER_SPECS))
def test_fp32_prefilter_never_changes_a_decision(case, mcig, monkeypatch):
    """The FP32 pre-filter of the accept test (device/mcig_device.cuh:accept_log) must give exactly the decisions of the plain
    FP64 test u <= exp(d): same seed, same streams => bit-identical per-walker averages, positions and acceptance counts
    (register and shared-memory kernels, full and selective acceptance paths, 32- and 53-bit uniforms)."""
    spec = PREFILTER_SPECS[case]
    res = []
    for flag in ("1", "0"):
        monkeypatch.setenv("MCIG_JIT_DEFINES", "MCIG_ACCEPT_PREFILTER=" + flag)
        for mode in (0, 1):
            mci = build_mci(mcig, spec, nwalkers=2048, mode=mode)
            assert ("MCIG_ACCEPT_PREFILTER " + flag) in mci.kernelSource()
            mci.integrate(spec["nmc"], False, False)
            wavg, _ = mci.walkerResults()
            res.append((flag, mode, wavg.copy(), mci.getAcceptanceRate(), np.array([mci.getX(walker=w) for w in (0, 1, 2047)])))
    for mode in (0, 1):
        a = [r for r in res if r[0] == "1" and r[1] == mode][0]
        b = [r for r in res if r[0] == "0" and r[1] == mode][0]
        assert np.array_equal(a[2], b[2]) and a[3] == b[3] and np.array_equal(a[4], b[4])
        assert 0.2 < a[3] < 0.8


ALL_VPO_SPECS = {
    "all32_gauss": dict(ndim=32, seed=2031, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_XND, 20, 1)], nmc=4000, steps=(0.2,)),
    "all24_expnd": dict(ndim=24, seed=2032, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 0, 1)], nmc=4000, steps=(0.6,)),
    # (starts in the typical set, sum |x_j| ~ ndim: from the mode at the origin an all-move of 96 coordinates is accepted with p ~ 1e-6)
    "all96_expnd_unrolled": dict(ndim=96, seed=2033, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 20, 1)], nmc=1000, steps=(0.15,),
                                 x0=[1.0 if j % 2 == 0 else -1.0 for j in range(96)]),
}


@pytest.mark.parametrize("case", sorted(ALL_VPO_SPECS))
def test_all_move_proto_views_equal_proto_arrays(case, mcig, monkeypatch):
    """All-moves in state memory read both sides of the acceptance test through ProtoViews over x and over the proposal when the
    sampling function is element-wise with protoElement (no proto-value arrays: device/mcig_device.cuh walk_state, MS_VPO). The
    recomputed proto values are the stored ones bit for bit, so per-walker results must not change (MCIG_ALL_VPO=0 keeps the arrays)."""
    spec = ALL_VPO_SPECS[case]
    res = {}
    for flag in ("1", "0"):
        monkeypatch.setenv("MCIG_ALL_VPO", flag)
        mci = build_mci(mcig, spec, nwalkers=1024, mode=0, placement=1)  # (the automatic choice at ndim 96 would be lane-split walkers)
        assert ("MS_MAIN_VPO = " + ("true" if flag == "1" else "false")) in mci.kernelSource()
        avg, err = mci.integrate(spec["nmc"], False, False)
        wavg, _ = mci.walkerResults()
        res[flag] = (np.array(avg), np.array(err), wavg.copy(), mci.getAcceptanceRate(), np.array([mci.getX(walker=w) for w in (0, 511, 1023)]))
    a, b = res["1"], res["0"]
    assert np.array_equal(a[2], b[2]) and a[3] == b[3] and np.array_equal(a[4], b[4])
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])
    assert 0.15 < a[3] < 0.85


def test_two_sampling_functions_multiply(mcig):
    """SamplingFunctionContainer multiplies the acceptances of all pdfs (src/SamplingFunctionContainer.cpp:41-48):
    exp(-r^2) * exp(-r^2) = exp(-2 r^2)  =>  <x^2> = 1/4 per coordinate. Also exercises the log-acceptance sum of the pre-filter."""
    mci = mcig.MCI(3)
    mci.setRngMode(0)
    mci.setSeed(11)
    mci.setNWalkers(1000)  # not a multiple of the block size
    mci.addSamplingFunction(mcig.ThreeDimGaussianPDF())
    mci.addSamplingFunction(mcig.Gauss(3))
    mci.addObservable(mcig.XYZSquared(), 0, 1)
    mci.setMRT2Step(0.8)
    mci.integrate(2000, False, False)
    avg, _ = mci.integrate(20000, False, False)
    cw = mci.crossWalkerError()
    assert np.all(np.abs(avg - 0.25) < 4.5*cw), (avg, cw)


def test_walker_count_not_multiple_of_warp(mcig):
    """Per-walker streams do not depend on how many walkers share the launch: walkers 0..32 of a 33-walker job equal those of a 1000-walker job."""
    spec = dict(ndim=3, seed=5, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 0, 1)], nmc=3000, steps=(1.0,))
    a = build_mci(mcig, spec, nwalkers=33, mode=0)
    b = build_mci(mcig, spec, nwalkers=1000, mode=0)
    a.integrate(3000, False, False)
    b.integrate(3000, False, False)
    assert np.array_equal(a.walkerResults()[0][0], b.walkerResults()[0][0, :33])


def test_box_muller_proposals_are_standard_normal(mcig):
    """Philox-mode Gaussian proposals: with no sampling function every proposal... is not available (no-pdf mode samples the box),
    so check through the chain instead: a Gaussian all-move of step s on a FLAT direction. Use ExpND pdf in 1-D with a huge box:
    simpler and sharper — the calibrated acceptance of a 1-D standard normal target under N(0, s^2) proposals is known:
    acc(s) = (2/pi) * atan(2/s). For s = 2: 0.5."""
    mci = mcig.MCI(1)
    mci.setRngMode(0)
    mci.setSeed(77)
    mci.setNWalkers(4096)
    mci.setTrialMove(mcig.SRRDType.Gaussian)
    mci.addSamplingFunction(mcig.Gauss(1))  # exp(-x^2): sigma^2 = 1/2  => in units of sigma the step below is 2*sqrt(2)*0.5
    mci.addObservable(mcig.X2(1), 0, 1)
    mci.setMRT2Step(np.sqrt(2.0))            # proposal sigma = 2 target sigmas
    mci.integrate(500, False, False)
    avg, _ = mci.integrate(4000, False, False)
    assert abs(avg[0] - 0.5) < 5*mci.crossWalkerError()[0]
    assert abs(mci.getAcceptanceRate() - (2/np.pi)*np.arctan(2/2.0)) < 2e-3


@pytest.mark.parametrize("srrd", list(range(10)))
def test_every_srrd_distribution_samples_the_target(srrd, mcig):
    """Any symmetric proposal leaves the target invariant: <x_i^2> = 1/2 for exp(-r^2) with all ten SRRDType proposals
    (closed-form device samplers in Philox mode), all-move and single-vector move."""
    for veclen in (0, 1):
        mci = mcig.MCI(3)
        mci.setRngMode(0)
        mci.setSeed(300 + srrd)
        mci.setNWalkers(2048)
        mci.setTrialMove(mcig.SRRDType(srrd), veclen)
        mci.addSamplingFunction(mcig.Gauss(3))
        mci.addObservable(mcig.XYZSquared(), 0, 1)
        mci.setMRT2Step(0.5)
        mci.integrate(3000, False, False)
        avg, _ = mci.integrate(6000, False, False)
        cw = mci.crossWalkerError()
        assert 0.05 < mci.getAcceptanceRate() < 0.98
        assert np.all(np.abs(avg - 0.5) < 5*cw), (srrd, veclen, avg, cw)


FLAT_PDF_SRC = """struct FlatPDF { static constexpr int NPAR = 0; static constexpr bool HAS_UPDATE = false; static constexpr bool ELEMENTWISE = false; const double * par;
  template <class X, class P> __device__ void protoFunction(const X &, P & pv) const { pv[0] = 0.; }
  template <class P> __device__ double samplingFunction(const P &) const { return 1.; }
  template <class PO, class PN> __device__ double acceptanceFunction(const PO &, const PN &) const { return 1.; } };"""

PARAM_LAWS = [  # (SRRDType, parameters, scipy law of the value (two-sided) or of its magnitude (symmetrised positive laws))
    (1, (0.5,), "norm", dict(scale=0.5), False), (2, (2.0,), "t", dict(df=2.0), False), (2, (5.5,), "t", dict(df=5.5), False),
    (3, (0.25,), "cauchy", dict(scale=0.25), False), (4, (2.5,), "expon", dict(scale=1/2.5), True), (5, (2.5, 0.5), "gamma", dict(a=2.5, scale=0.5), True),
    (5, (3.0, 2.0), "gamma", dict(a=3.0, scale=2.0), True), (6, (1.5, 0.8), "weibull_min", dict(c=1.5, scale=0.8), True),
    (7, (-0.5, 0.6), "lognorm", dict(s=0.6, scale=np.exp(-0.5)), True), (8, (3.0,), "chi2", dict(df=3.0), True), (9, (4.0, 6.0), "f", dict(dfn=4.0, dfd=6.0), True),
    # the default-parameter closed forms, through the same check
    (2, (), "t", dict(df=1.0), False), (4, (), "expon", dict(), True), (7, (), "lognorm", dict(s=1.0), True), (8, (), "chi2", dict(df=1.0), True),
    (9, (), "f", dict(dfn=1.0, dfd=1.0), True),
    # shapes without a closed form: Marsaglia & Tsang's test with a fixed number of tries (srrd_gamma_mt)
    (5, (2.3, 0.5), "gamma", dict(a=2.3, scale=0.5), True), (5, (0.3, 1.0), "gamma", dict(a=0.3), True), (5, (1.0001, 1.0), "gamma", dict(a=1.0001), True),
    (5, (100.25, 0.01), "gamma", dict(a=100.25, scale=0.01), True), (8, (2.5,), "chi2", dict(df=2.5), True), (9, (3.0, 200.0), "f", dict(dfn=3.0, dfd=200.0), True),
    (9, (2.7, 4.1), "f", dict(dfn=2.7, dfd=4.1), True),
]


@pytest.mark.parametrize("srrd,par,law,kw,positive", PARAM_LAWS)
def test_parameterised_proposals_follow_their_law(srrd, par, law, kw, positive, mcig):
    """Moves built around a pre-made distribution (mcig_set_srrd_params; reference: the rdist constructor argument, include/mci/SRRDAllMove.hpp:45-58):
    the Philox-mode closed forms must produce exactly that law. Under a flat sampling function every proposal is accepted, so the increments of the
    stored positions ARE step x draw: Kolmogorov-Smirnov against scipy's law (two-sided laws directly; symmetrised positive laws by magnitude, plus a
    fair sign)."""
    from scipy import stats
    mcig.register_plugin(0, "FlatPDF", "FlatPDF", FLAT_PDF_SRC, ndim=0, nvalues=1)
    mci = mcig.MCI(1)
    mci.setRngMode(0)
    mci.setSeed(4000 + 10*srrd + len(par))
    mci.setNWalkers(64)
    mci.setTrialMove(mcig.SRRDType(srrd), 0, params=par or None)
    mci.setMRT2Step(1.0)
    mci.addSamplingFunction(mcig.SamplingFunction("FlatPDF"))
    mci.addObservable(mcig.XND(1), 1, 1, False, mcig.EstimatorType.Noop)
    n = 8192
    mci.integrate(n, False, False)
    assert mci.getAcceptanceRate() == 1.0
    d = np.concatenate([np.diff(mci.obsData(0, walker=w, nobs=1)[:, 0]) for w in range(0, 64, 4)])
    # heavy tails: the walker drifts far from 0 and the increments lose low bits; nothing a distribution test sees
    if positive:
        assert abs(np.mean(d > 0) - 0.5) < 5*0.5/np.sqrt(d.size)
        d = np.abs(d)
    ks = stats.kstest(d, getattr(stats, law)(**kw).cdf)
    assert ks.pvalue > 1e-4, (srrd, par, ks)


def test_parameterised_proposals_argument_checks(mcig):
    from mcintegratorplusplus_b200._capi import McigError
    mci = mcig.MCI(2)
    mci.setRngMode(0)
    mci.addSamplingFunction(mcig.Gauss(2))
    mci.addObservable(mcig.XND(2), 0, 1)
    with pytest.raises(McigError, match="takes 2 parameter"):
        mci.setTrialMove(mcig.SRRDType.Gamma, 0, params=(2.0,))
    with pytest.raises(McigError, match="positive"):
        mci.setTrialMove(mcig.SRRDType.Student, 0, params=(-1.0,))
    mci.setTrialMove(mcig.SRRDType.Gamma, 0, params=(2.3, 1.0))  # no closed form: fixed-count Marsaglia-Tsang in the Philox modes
    mci.prebuild()
    mci.setRngMode(2)  # replay mode consumes the libstdc++ outputs of any parameter set
    mci.setTrialMove(mcig.SRRDType.Gamma, 0, params=(2.3, 1.0))
    mci.prebuild()


@pytest.mark.parametrize("ndim,placement", [(3, None), (3, 1), (24, None), (80, None)])
def test_user_defined_move_keeps_detailed_balance(ndim, placement, mcig):
    """A user-defined trial move (device functor of kind MCIG_PLUGIN_MOVE; the reference's TrialMoveInterface subclasses): a DRIFTED uniform proposal is
    only correct together with its acceptance factor (0 where the reverse move is impossible). <x_i> = 0 and <x_i^2> = 1/2 under exp(-r^2) on the
    register, shared-memory and block-by-block draw paths; the same move with the factor forced to 1 is visibly biased."""
    from prod import register_test_plugins
    register_test_plugins(mcig)
    biased = """template <int NDIM> struct DriftNoFactor { static constexpr int NPAR = 1; const double * par;
      template <class XO, class XN, class T, class U> __device__ double trialMove(const XO & xo, XN & xn, const double * st, T ty, const U & u) const {
        for (int i = 0; i < NDIM; ++i) { xn[i] = xo[i] + st[ty.of(i)]*(2.*u(i) - 1.) + par[0]; } return 1.; } };"""
    mcig.register_plugin(4, "DriftNoFactor", "DriftNoFactor<{ndim}>", biased, ndim=0, nvalues=0, npar=1)
    out = {}
    for name in ("DriftMove", "DriftNoFactor"):
        mci = mcig.MCI(ndim)
        mci.setRngMode(0)
        mci.setSeed(99)
        mci.setNWalkers(4096)
        step = 1.6/ndim**0.5
        mci.setTrialMove(mcig.Move(name, (0.6*step/ndim,)))  # (the reverse move is impossible with probability drift/step per coordinate)
        mci.setMRT2Step(step)
        if placement is not None:
            mci.setStatePlacement(placement)
        mci.addSamplingFunction(mcig.Gauss(ndim))
        mci.addObservable(mcig.XND(ndim), 0, 1)
        mci.addObservable(mcig.X2(ndim), 0, 1)
        mci.integrate(2000, False, False)
        avg, _ = mci.integrate(4000, False, False)
        out[name] = (avg, mci.crossWalkerError(), mci.getAcceptanceRate())
    avg, cw, acc = out["DriftMove"]
    assert 0.05 < acc < 0.9
    assert np.all(np.abs(avg[:ndim]) < 5*cw[:ndim]) and np.all(np.abs(avg[ndim:] - 0.5) < 5*cw[ndim:]), (avg, cw)
    bavg, bcw, _ = out["DriftNoFactor"]
    if ndim == 3:
        assert np.mean(bavg[:ndim]) > 10*np.mean(bcw[:ndim])  # the drift shows when the factor is dropped


def test_philox_rounds_option(mcig):
    """Philox4x32-7 (opt-in) is a different, still valid stream: same expectation, different per-walker values."""
    from mcintegratorplusplus_b200._capi import McigError
    spec = dict(ndim=3, seed=1, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 0, 1)], nmc=4000, steps=(1.0,))
    out = {}
    for r in (10, 7):
        mci = build_mci(mcig, spec, nwalkers=4096, mode=0)
        mci.setPhiloxRounds(r)
        mci.integrate(1000, False, False)
        avg, _ = mci.integrate(4000, False, False)
        assert abs(avg[0] - 0.5) < 5*mci.crossWalkerError()[0]
        out[r] = mci.walkerResults()[0].copy()
    assert not np.array_equal(out[10], out[7])
    with pytest.raises(McigError):
        build_mci(mcig, spec).setPhiloxRounds(3)


@pytest.mark.parametrize("name", ["auto_default", "vec3_types", "ms_sub_ut5", "ut4_fixed"])
def test_device_resident_calibration_equals_host_loop(name, mcig):
    """findMRT2Step with the feedback rule applied by a controller kernel (no host round trip per iteration) must follow exactly
    the trajectory of the host loop: same calibrated step sizes, same number of iterations, same Philox cursor afterwards
    (hence bit-identical results of the following run)."""
    spec = dict(configs.RUNS[name])
    out = []
    for on in (1, 0):
        mci = build_mci(mcig, spec, nwalkers=2048, mode=0)
        mci.setDeviceCalibration(on)
        avg, err = mci.integrate(12288, True, False)
        nt = max(1, spec.get("ntypes", 1))
        out.append(([mci.getMRT2Step(i) for i in range(nt)], mci.getCalibrationIterations(), avg.copy(), err.copy(), mci.getAcceptanceRate()))
    assert out[0][0] == out[1][0] and out[0][1] == out[1][1] and out[0][1] >= 1
    assert np.array_equal(out[0][2], out[1][2]) and np.array_equal(out[0][3], out[1][3]) and out[0][4] == out[1][4]


@pytest.mark.parametrize("name", ["vec_exp4", "ndim_vec16", "ndim_vec64_v4", "vec_ortho_types", "ndim_vec300_v3_types"])
def test_lazy_accumulation_equals_per_step_accumulation(name, mcig):
    """Element-wise observables under single-vector moves are accumulated as value x dwell time; the trajectories are untouched
    (same acceptance, same final positions) and the sums agree with the per-step additions to rounding."""
    spec = dict(configs.RUNS[name])
    spec["obs"] = [(orc.OBS_XND, 0, 1), (orc.OBS_X2, 16, 1), (orc.OBS_X2SUM, 1, 1)]
    out = []
    for lazy in (1, 0):
        mci = build_mci(mcig, spec, nwalkers=64, mode=0)
        mci.setLazyAccumulation(lazy)
        avg, err = mci.integrate(4096, False, False)
        out.append((avg.copy(), err.copy(), mci.getAcceptanceRate(), list(mci.getX()), "LazyAccu" in mci.kernelSource()))
    assert out[0][4] and not out[1][4]
    assert out[0][2] == out[1][2] and out[0][3] == out[1][3]
    # zero-mean components: the sums of |x| ~ 1 values differ by rounding only, i.e. by ~1e-16 relative to the typical magnitude
    assert np.max(np.abs(out[0][0] - out[1][0])) < 1e-13, np.max(np.abs(out[0][0] - out[1][0]))
    assert np.allclose(out[0][1], out[1][1], rtol=1e-9, atol=1e-16)


@pytest.mark.parametrize("name", ["c1_simple_short", "vec_exp4", "nopdf_box", "mixed"])
def test_stream_position_is_resumable_across_a_2_32_boundary(name, mcig):
    """The Philox group counter is 64 bit and random-access: started 1000 groups before a 2^32 boundary, one run of 2n steps equals two
    runs of n steps, on the register and on the shared-memory walkers alike, and the stream position advances by the groups consumed."""
    spec = dict(configs.RUNS[name])
    n = 4000
    start = 3*2**32 - 1000
    out = []
    # (0, 1, 1): the dynamically scheduled register kernel, whose inner loop runs on a split (32-bit counter, constant high word) form of the
    # group counter and therefore has to end its chunks exactly where the low word wraps
    for placement, pieces, dyn in ((0, 1, 0), (0, 2, 0), (1, 1, 0), (0, 1, 1)):
        if placement == 1 and spec["pdf_id"] == orc.PDF_NONE:
            continue
        mci = build_mci(mcig, spec, nwalkers=96, mode=0, placement=placement)
        mci.setDynamicScheduling(dyn)
        mci.setStreamPosition(start)
        for _ in range(pieces):
            mci.integrate(n//pieces, False, False)
        groups_per_step = 1
        assert mci.getStreamPosition() == start + n*groups_per_step
        out.append((mci.getAcceptanceRate() if pieces == 1 else None, list(mci.getX())))
    assert all(o[1] == out[0][1] for o in out)
    assert all(o[0] == out[0][0] for o in out if o[0] is not None)


def test_warp_specialised_kernel_reproduces_the_single_role_kernel():
    """The experimental producer / consumer kernel (MCIG_WS=1: Philox blocks through a shared-memory ring with mbarrier hand-over) consumes the same
    blocks in the same order: positions, acceptance and sums are bit-identical to the default kernel, also across a 2^32 boundary of the group
    counter. The switch is read once per process, hence the subprocesses."""
    import json
    import os
    import subprocess
    import sys
    code = r'''
import json, sys
sys.path.insert(0, %r)
import mcintegratorplusplus_b200 as m
out = []
for start in (0, 5*2**32 - 3000):
    mci = m.MCI(3); mci.setRngMode(0); mci.setSeed(99); mci.setNWalkers(256)
    mci.addSamplingFunction(m.ThreeDimGaussianPDF()); mci.addObservable(m.XSquared(), 0, 1); mci.setMRT2Step(1.0)
    mci.setDynamicScheduling(1); mci.setStreamPosition(start)
    avg, err = mci.integrate(20000, False, False)
    out.append([float(avg[0]).hex(), mci.getAcceptanceRate(), [float(v).hex() for v in mci.getX()], "mcig_walk_ws" in mci.kernelSource()])
print(json.dumps(out))
''' % os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    res = []
    for ws in ("0", "1"):
        env = dict(os.environ, MCIG_WS=ws)
        r = subprocess.run([sys.executable, "-c", code], env=env, capture_output=True, text=True, timeout=600)
        assert r.returncode == 0, r.stderr[-2000:]
        res.append(json.loads(r.stdout.strip().splitlines()[-1]))
    assert [o[3] for o in res[0]] == [False, False] and [o[3] for o in res[1]] == [True, True]
    assert [o[:3] for o in res[0]] == [o[:3] for o in res[1]]
