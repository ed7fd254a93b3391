"""GPU: the CUDA walk path (through the C-ABI) in REPLAY mode against the C oracle and the reference's golden vectors.

Replay mode feeds the kernel the same std::mt19937_64 / libstdc++ distribution outputs the reference consumes, so:
  * accept/reject sequences, acceptance counts, final positions and calibrated step sizes must be BIT-EXACT
    (the only non-IEEE-pinned operation on the path is exp(): CUDA vs glibc may differ by 1 ulp, which flips a decision
    only if the uniform lands inside that ulp, p ~ 1e-16 per step — a flip would show up here, not be hidden);
  * accumulator sums / averages must agree to 1e-12 relative (BASELINE.json north_star);
  * estimator errors (MJ / FC / uncorrelated) to 1e-9 relative.
"""
import ctypes as C

import numpy as np
import pytest

import configs
import orc
from conftest import fromhex
from prod import build_mci

pytestmark = pytest.mark.gpu

AVG_RTOL, ERR_RTOL = 1e-12, 1e-9


def _close(a, b, rtol, atol=1e-15):
    a, b = np.asarray(a, dtype=float), np.asarray(b, dtype=float)
    return np.all(np.abs(a - b) <= atol + rtol*np.abs(b))


@pytest.mark.parametrize("name", sorted(configs.RUNS))
def test_replay_vs_oracle_and_golden(name, mcig, oracle, golden_runs):
    spec = configs.RUNS[name]
    g = golden_runs[name]
    if configs.in_oracle(name):
        ref = oracle.run(configs.make(name))
        assert ref["avg"] == fromhex(g["avg"])  # oracle itself is pinned to the reference
    else:  # distributions the C oracle does not restate: the reference's own outputs (goldens) are the checker
        ref = {"avg": fromhex(g["avg"]), "err": fromhex(g["err"]), "acc_rate": float.fromhex(g["acc_rate"]), "x_final": fromhex(g["x_final"]),
               "steps_final": fromhex(g["steps_final"]), "n_acc": g["n_acc"]}
    mci = build_mci(mcig, spec)
    avg, err = mci.integrate(spec["nmc"], spec.get("do_find", False), spec.get("do_decorr", False))
    # bit-exact integer / control-flow quantities
    n = spec["nmc"]
    if n > 0:
        assert round(mci.getAcceptanceRate()*n) == ref["n_acc"], "accept count differs"
    assert mci.getAcceptanceRate() == ref["acc_rate"]
    assert list(mci.getX()) == ref["x_final"], "final position not bit-exact"
    nt = max(1, spec.get("ntypes", 1))
    assert [mci.getMRT2Step(i) for i in range(nt)] == ref["steps_final"], "calibrated step sizes not bit-exact"
    assert _close(avg, ref["avg"], AVG_RTOL), (avg, ref["avg"])
    assert _close(err, ref["err"], ERR_RTOL, atol=1e-18), (err, ref["err"])


USER_PERIODIC_SRC = """template <int NDIM> struct UserPeriodic { static constexpr int NPAR = 2*NDIM; const double * par; // lb[NDIM], ub[NDIM]
  __device__ void wrap(int i, double & x) const { const double l = par[i], u = par[NDIM + i]; while (x < l) { x += u - l; } while (x > u) { x -= u - l; } }
  __device__ double scale(int i, double u01) const { return par[i] + u01*(par[NDIM + i] - par[i]); } };"""


@pytest.mark.parametrize("placement", [None, 1, 2])
@pytest.mark.parametrize("name", ["ut2_irange", "vec_ortho_types", "ms_ortho", "nopdf_box", "exbasic_1", "gauss_vec6_v3", "par_fisher_all"])
def test_user_defined_domain_functor_reproduces_the_builtin_periodic_domain(name, placement, mcig, golden_runs):
    """MCI::setDomain with a user-defined DomainInterface (device functor of plugin kind MCIG_PLUGIN_DOMAIN): a functor restating the periodic
    wrap (src/OrthoPeriodicDomain.cpp:38-68) must reproduce the reference's goldens of the setIRange runs bit for bit -- all-, vec- and
    MultiStep moves, plain sampling of the box without sampling function (scaleToDomain + volume), calibration with the step capped by the
    domain size -- on every state placement."""
    spec = configs.RUNS[name]
    g = golden_runs[name]
    ndim = spec["ndim"]
    lb = np.full(ndim, spec["lb"], dtype=float) if np.isscalar(spec["lb"]) else np.asarray(spec["lb"], dtype=float)
    ub = np.full(ndim, spec["ub"], dtype=float) if np.isscalar(spec["ub"]) else np.asarray(spec["ub"], dtype=float)
    mcig.register_plugin(3, "UserPeriodic", "UserPeriodic<{ndim}>", USER_PERIODIC_SRC, ndim=0, nvalues=0, npar=0)
    mci = build_mci(mcig, spec, placement=placement)
    x0 = mci.getX()
    with pytest.raises(Exception, match="number of functor parameters"):
        mci.setDomain(mcig.Domain("UserPeriodic", list(lb)), ub - lb, float(np.prod(ub - lb)))
    mcig.register_plugin(3, "UserPeriodic%d" % ndim, "UserPeriodic<{ndim}>", USER_PERIODIC_SRC, ndim=ndim, nvalues=0, npar=2*ndim)
    mci.setDomain(mcig.Domain("UserPeriodic%d" % ndim, list(lb) + list(ub)), ub - lb, float(np.prod(ub - lb)))
    mci.setX(x0)  # (build_mci's setIRange already wrapped the start position; a user domain takes positions as given)
    avg, err = mci.integrate(spec["nmc"], spec.get("do_find", False), spec.get("do_decorr", False))
    assert mci.getAcceptanceRate() == float.fromhex(g["acc_rate"])
    assert list(mci.getX()) == fromhex(g["x_final"])
    nt = max(1, spec.get("ntypes", 1))
    assert [mci.getMRT2Step(i) for i in range(nt)] == fromhex(g["steps_final"])
    assert _close(avg, fromhex(g["avg"]), AVG_RTOL) and _close(err, fromhex(g["err"]), ERR_RTOL, atol=1e-18)


@pytest.mark.parametrize("placement", [0, 1, 2])
@pytest.mark.parametrize("name", ["c1_simple_short", "vec_exp4", "ms_default4", "ms_sub_ut5", "all_types", "vec3_types"])
def test_register_and_smem_paths_agree(name, placement, mcig, oracle):
    """All state placements (registers / shared memory / global memory) must reproduce the oracle."""
    spec = configs.RUNS[name]
    ref = oracle.run(configs.make(name))
    mci = build_mci(mcig, spec, placement=placement)
    avg, err = mci.integrate(spec["nmc"], spec.get("do_find", False), spec.get("do_decorr", False))
    assert mci.getAcceptanceRate() == ref["acc_rate"]
    assert list(mci.getX()) == ref["x_final"]
    assert _close(avg, ref["avg"], AVG_RTOL) and _close(err, ref["err"], ERR_RTOL, atol=1e-18)


@pytest.mark.parametrize("placement", [0, 1, 2])
@pytest.mark.parametrize("name", configs.C3_SHAPES)
def test_c3_shapes_replay_on_every_placement(name, placement, mcig, oracle, golden_runs):
    """The shapes BASELINE configs[2] runs (ExpNDPDF / Gauss + XND, Block(20), ndim 32 and 64: MultiStepMove with and without sub-pdf,
    all-move, single-index vec move) are the register- / footprint-limited kernels where a spill or unroll bug would hide: every
    placement (registers with local-memory indexing, shared memory, global memory) replays the reference bit for bit.
    Reference: src/MultiStepMove.cpp:6-47, benchmark/bench_throughput_ndim_single/main.cpp:26-50."""
    spec = configs.RUNS[name]
    g = golden_runs[name]
    ref = oracle.run(configs.make(name))
    assert ref["avg"] == fromhex(g["avg"]) and ref["x_final"] == fromhex(g["x_final"])
    mci = build_mci(mcig, spec, placement=placement)
    avg, err = mci.integrate(spec["nmc"], False, False)
    assert round(mci.getAcceptanceRate()*spec["nmc"]) == g["n_acc"], "accept count differs"
    assert list(mci.getX()) == ref["x_final"], "final position not bit-exact"
    assert _close(avg, ref["avg"], AVG_RTOL), np.max(np.abs(avg - np.array(ref["avg"])))
    assert _close(err, ref["err"], ERR_RTOL, atol=1e-18)


@pytest.mark.parametrize("name", sorted(configs.LANE_SPLIT))
def test_lane_split_walkers_replay_bit_exact(name, mcig, oracle, golden_runs):
    """State placement 3: one walker spread over several lanes of a warp (all-moves over many coordinates, src of the move:
    include/mci/SRRDAllMove.hpp:67-80). In replay mode the two sums of the acceptance exp(sum po - sum pn) (test/common/TestMCIFunctions.hpp:
    173-177, 240-245) travel through the lanes in coordinate order, so every accept decision, the trajectory and the calibrated step sizes are
    the reference's bit for bit; the accumulators are per coordinate and keep the reference's order."""
    spec = configs.RUNS[name]
    g = golden_runs[name]
    mci = build_mci(mcig, spec, placement=3)
    avg, err = mci.integrate(spec["nmc"], spec.get("do_find", False), spec.get("do_decorr", False))
    src = mci.kernelSource()
    assert "walk_kernel_lanes" in src and "LANES = %d" % configs.LANE_SPLIT[name] in src
    assert mci.getAcceptanceRate() == float.fromhex(g["acc_rate"])
    assert list(mci.getX()) == fromhex(g["x_final"]), "final position not bit-exact"
    nt = max(1, spec.get("ntypes", 1))
    assert [mci.getMRT2Step(i) for i in range(nt)] == fromhex(g["steps_final"])
    assert _close(avg, fromhex(g["avg"]), AVG_RTOL), np.max(np.abs(avg - np.array(fromhex(g["avg"]))))
    assert _close(err, fromhex(g["err"]), ERR_RTOL, atol=1e-18)


@pytest.mark.parametrize("name", ["ndim_all64_b20", "lanes_all128_auto"])
def test_lane_split_walkers_philox_equal_shared_memory_walkers(name, mcig):
    """Production mode: the lanes draw from the same (group, walker, block) -> word mapping as every other placement, so a lane-split walker
    follows the same chain as its shared-memory twin; only the association of the log-acceptance sum differs (butterfly over the lanes), which
    can flip a decision only when the uniform lies within an ulp of the acceptance."""
    spec = dict(configs.RUNS[name], do_find=False, do_decorr=False)
    out = []
    for placement in (1, 3):
        mci = build_mci(mcig, spec, nwalkers=192, mode=0, placement=placement)
        avg, err = mci.integrate(1000, False, False)
        out.append((avg.copy(), err.copy(), mci.getAcceptanceRate(), np.array([mci.getX(walker=w) for w in (0, 100, 191)])))
    assert out[0][2] == out[1][2], "accept counts differ"
    assert np.array_equal(out[0][3], out[1][3]), "positions differ"
    scale = max(1.0, float(np.max(np.abs(out[0][0]))))
    assert np.max(np.abs(out[0][0] - out[1][0])) <= 1e-12*scale
    assert np.allclose(out[0][1], out[1][1], rtol=1e-9, atol=1e-15)


@pytest.mark.parametrize("name", configs.C3_SHAPES)
def test_c3_shapes_philox_placements_agree(name, mcig):
    """Production (Philox) kernels of the same shapes: the shared- and global-memory placements run the same code over different views
    and must agree bit for bit; the register placement may contract products differently, so it must reproduce every accept decision
    (positions and acceptance count bit-identical) and the averages to rounding."""
    spec = configs.RUNS[name]
    n = min(spec["nmc"], 2000)
    out = []
    for placement in (0, 1, 2):  # (lane-split walkers, which the automatic choice would pick for some of these shapes: tests above)
        mci = build_mci(mcig, spec, nwalkers=160, mode=0, placement=placement)
        mci.setLazyAccumulation(0)
        avg, err = mci.integrate(n, False, False)
        out.append((avg.copy(), err.copy(), mci.getAcceptanceRate(), [list(mci.getX(walker=w)) for w in (0, 77, 159)]))
    assert np.array_equal(out[1][0], out[2][0]) and np.array_equal(out[1][1], out[2][1]) and out[1][2:] == out[2][2:]
    assert out[0][2] == out[1][2]
    assert np.allclose(out[0][3], out[1][3], rtol=1e-12, atol=1e-14)
    scale = max(1.0, float(np.max(np.abs(out[1][0]))))
    assert np.max(np.abs(out[0][0] - out[1][0])) <= 1e-12*scale


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("ndim,veclen,nsub", [(6, 1, 0), (4, 1, 0), (5, 1, 7), (6, 3, 5), (24, 1, 0), (3, 1, 1)])
def test_multistep_quad_draw_groups_keep_the_stream_properties(ndim, veclen, nsub, mode, mcig):
    """MultiStepMove in the Philox modes packs four sub-steps into one draw group (3 Philox blocks instead of 4) and the outer accept uniform into the
    group of the last nsteps % 4 sub-steps (device/mcig_device.cuh: MCIG_MS_QUADS). The stream properties must survive for every sub-step count and
    vector length: two launches continue one launch's streams, the dynamically scheduled kernel equals the static one, a shard of the walkers
    reproduces its walkers of the full job, and the state placements agree (shared = global bit for bit, registers to the accept decisions)."""
    spec = dict(ndim=ndim, seed=77, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_XND, 0, 1)], nmc=1200, move_type=orc.MOVE_MULTISTEP, veclen=veclen,
                ms_nsteps=nsub, ms_sub_pdf_id=orc.PDF_EXPND, steps=(0.6,), x0=[0.1*(-1)**j for j in range(ndim)])

    def run(placement=None, parts=(1200,), dyn=-1, shard=None):
        mci = build_mci(mcig, spec, nwalkers=64 if shard else 192, mode=mode, placement=placement)
        if shard:
            mci.setNWalkers(64, global_offset=64, total=192)
        mci.setLazyAccumulation(0)
        mci.setDynamicScheduling(dyn)
        for n in parts:
            avg, err = mci.integrate(n, False, False)
        W = 64 if shard else 192
        return avg.copy(), mci.getAcceptanceRate(), [list(mci.getX(walker=w)) for w in (0, 33, W - 1)]

    base = run()
    # the same streams whatever the launch chunking (400 + 800 steps: not multiples of the sub-step count or of four)
    assert run(parts=(400, 800))[2] == base[2]
    assert run(parts=(1, 1199))[2] == base[2]
    # dynamic chunk scheduling (register kernels): same numbers
    if ndim <= 8:
        d = run(dyn=1)
        assert d[2] == base[2] and d[1] == base[1]
    # a shard reproduces its walkers
    sh = run(shard=True)
    full = build_mci(mcig, spec, nwalkers=192, mode=mode)
    full.setLazyAccumulation(0)
    full.integrate(1200, False, False)
    assert sh[2][0] == list(full.getX(walker=64)) and sh[2][2] == list(full.getX(walker=127))
    # placements
    outs = [run(placement=pl) for pl in (0, 1, 2)]
    assert outs[1][2] == outs[2][2] and outs[1][1] == outs[2][1] and np.array_equal(outs[1][0], outs[2][0])
    assert outs[0][1] == outs[1][1]
    assert np.allclose(outs[0][2], outs[1][2], rtol=1e-12, atol=1e-14)


def test_accept_sequence_and_trajectory_bit_exact(mcig, oracle):
    """Every step: positions from a Full XND accumulator vs the trajectory implied by the oracle's draws + accept bits."""
    nmc = 20000
    spec = dict(ndim=3, seed=5649871, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XND, 1, 1, False, orc.EST_UNCORRELATED)], nmc=nmc, steps=(1.0,),
                x0=(0.3, -0.2, 0.1))
    cfg = orc.make_config(3, 5649871, orc.PDF_GAUSS3D, spec["obs"], nmc, steps=(1.0,), x0=spec["x0"])
    tr = oracle.run(cfg, trace=True)
    draws = tr["draws"].reshape(nmc, 4)
    x = np.array(spec["x0"])
    traj = np.zeros((nmc, 3))
    for t in range(nmc):
        if tr["accepted"][t]:
            x = x + 1.0*draws[t, :3]
        traj[t] = x
    mci = build_mci(mcig, spec)
    mci.setKeepSamples(True)  # the uncorrelated estimator of a small observable runs inside the walk; keep the series for this check
    mci.integrate(nmc, False, False)
    data = mci.obsData(0, walker=0, nobs=3)
    assert data.shape == (nmc, 3)
    assert np.array_equal(data, traj), "trajectory differs from the reference's at step %d" % int(np.argmax((data != traj).any(axis=1)))
    prev = np.vstack([np.array(spec["x0"])[None, :], data[:-1]])
    acc_gpu = (data != prev).any(axis=1)
    assert np.array_equal(acc_gpu, tr["accepted"].astype(bool))


def test_vec_move_accept_sequence(mcig, oracle):
    nmc = 30000
    spec = dict(ndim=4, seed=1337, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 1, 1, False, orc.EST_UNCORRELATED)], nmc=nmc, move_type=orc.MOVE_VEC,
                veclen=1, steps=(3.0,), x0=[0.1, -0.05, 0.1, -0.05])
    cfg = orc.make_config(4, 1337, orc.PDF_EXPND, spec["obs"], nmc, move_type=orc.MOVE_VEC, veclen=1, steps=(3.0,), x0=spec["x0"])
    tr = oracle.run(cfg, trace=True)
    for placement in (0, 1):
        mci = build_mci(mcig, spec, placement=placement)
        mci.setKeepSamples(True)
        mci.integrate(nmc, False, False)
        data = mci.obsData(0, walker=0, nobs=4)
        prev = np.vstack([np.array(spec["x0"])[None, :], data[:-1]])
        assert np.array_equal((data != prev).any(axis=1), tr["accepted"].astype(bool))


def test_multi_walker_replay_equals_independent_ranks(mcig, oracle):
    """W walkers with per-walker seeds == W reference ranks (src/MPIMCI.cpp:83-92): per-walker results and the combination."""
    spec = configs.RUNS["full_mj"]
    cfg = configs.make("full_mj")
    seeds = [11, 22, 33, 44, 55]
    f = oracle.lib.mcio_run_ranks
    f.restype = C.c_int
    f.argtypes = [C.POINTER(orc.Config), C.POINTER(C.c_uint64), C.c_int, C.POINTER(orc.Result), C.POINTER(orc.Result), C.POINTER(orc.Trace)]
    comb = orc.Result()
    per = (orc.Result*len(seeds))()
    assert f(C.byref(cfg), (C.c_uint64*len(seeds))(*seeds), len(seeds), C.byref(comb), per, None) == 0
    mci = build_mci(mcig, spec, nwalkers=len(seeds), seeds=seeds)
    avg, err = mci.integrate(spec["nmc"], False, False)
    assert _close(avg, comb.avg[:4], AVG_RTOL) and _close(err, comb.err[:4], ERR_RTOL)
    wavg, werr = mci.walkerResults()
    for w in range(len(seeds)):
        assert _close(wavg[:, w], per[w].avg[:4], AVG_RTOL)
        assert _close(werr[:, w], per[w].err[:4], ERR_RTOL)
        assert list(mci.getX(walker=w)) == list(per[w].x_final[:3])
    assert mci.getAcceptanceRate() == comb.acc_rate


def test_multi_walker_auto_calibration_matches_mpi_semantics(mcig, oracle):
    """Calibration / decorrelation with rank-averaged acceptance rate and estimates (src/MCIntegrator.cpp:131-138, 203-230)."""
    spec = dict(configs.RUNS["auto_default"])
    cfg = configs.make("auto_default")
    seeds = [101, 202, 303, 404]
    cfg.nranks_for_minstat = len(seeds)  # MIN_STAT / MIN_NMC shrink with the number of ranks: src/MCIntegrator.cpp:107, :193
    f = oracle.lib.mcio_run_ranks
    f.restype = C.c_int
    f.argtypes = [C.POINTER(orc.Config), C.POINTER(C.c_uint64), C.c_int, C.POINTER(orc.Result), C.POINTER(orc.Result), C.POINTER(orc.Trace)]
    comb = orc.Result()
    assert f(C.byref(cfg), (C.c_uint64*len(seeds))(*seeds), len(seeds), C.byref(comb), None, None) == 0
    mci = build_mci(mcig, spec, nwalkers=len(seeds), seeds=seeds)
    avg, err = mci.integrate(spec["nmc"], True, True)
    # the rank-sum order of the rate is implementation-defined in MPI; step sizes agree to rounding, results to 1e-9
    assert mci.getMRT2Step(0) == pytest.approx(comb.steps_final[0], rel=1e-12)
    assert mci.getAcceptanceRate() == pytest.approx(comb.acc_rate, abs=2e-4)
    assert abs(avg[0] - comb.avg[0]) < 3*np.hypot(err[0], comb.err[0])
    assert abs(avg[0] - 0.5) < 3*err[0]


def test_state_persists_across_integrate_calls(mcig, oracle):
    """Walker position and RNG stream continue across integrate calls: two 2048-step calls == oracle twice."""
    spec = dict(configs.RUNS["c1_simple_short"])
    mci = build_mci(mcig, spec)
    a1, _ = mci.integrate(2048, False, False)
    a2, _ = mci.integrate(2048, False, False)
    cfg = configs.make("c1_simple_short")
    cfg.nmc = 2048
    r1 = oracle.run(cfg)
    # second leg: start from r1's final x with the generator advanced by 2048 steps' draws -> emulate with one 4096 run's accept count
    cfg.nmc = 4096
    r12 = oracle.run(cfg)
    assert _close(a1, r1["avg"], AVG_RTOL)
    assert list(mci.getX()) == r12["x_final"]
    assert _close(0.5*(a1[0] + a2[0]), r12["avg"][0], 1e-12)


def test_error_behaviour_on_device(mcig):
    from mcintegratorplusplus_b200._capi import McigError
    mci = build_mci(mcig, dict(configs.RUNS["c1_simple_short"], obs=[(orc.OBS_XSQUARED, 7, 1)]))
    with pytest.raises(McigError, match="not a multiple of the requested block size"):
        mci.integrate(100, False, False)  # src/BlockAccumulator.cpp:11-13
    mci = build_mci(mcig, dict(configs.RUNS["c1_simple_short"], obs=[(orc.OBS_XSQUARED, 1, 1, False, orc.EST_MJBLOCKER)]))
    with pytest.raises(McigError, match="power of two"):
        mci.integrate(1000, False, False)  # src/MJBlocker.cpp:28-30
    mci = build_mci(mcig, dict(configs.RUNS["c1_simple_short"], obs=[(orc.OBS_XSQUARED, 1, 1, False, orc.EST_FCBLOCKER)]))
    with pytest.raises(McigError, match=">= 50"):
        mci.integrate(40, False, False)  # src/Estimators.cpp:196-198


@pytest.mark.parametrize("name", ["c1_simple", "mixed", "block8_skip2_corr", "full_mj", "tp3g_small", "ms_default4", "all_types", "nopdf_box", "auto_default"])
def test_dynamic_chunk_scheduling_bit_exact(name, mcig, oracle):
    """The persistent, work-stealing variant of the register kernel cuts every chain into chunks that may run on different SMs;
    chain state travels through L2. It must reproduce the oracle exactly like the static kernel (replay mode)."""
    spec = configs.RUNS[name]
    ref = oracle.run(configs.make(name))
    mci = build_mci(mcig, spec, placement=0)
    mci.setDynamicScheduling(1)
    avg, err = mci.integrate(spec["nmc"], spec.get("do_find", False), spec.get("do_decorr", False))
    assert mci.getAcceptanceRate() == ref["acc_rate"]
    assert list(mci.getX()) == ref["x_final"]
    assert _close(avg, ref["avg"], AVG_RTOL) and _close(err, ref["err"], ERR_RTOL, atol=1e-18)


def test_dynamic_equals_static_in_philox_mode(mcig):
    spec = dict(ndim=3, seed=4, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 0, 1), (orc.OBS_XYZSQUARED, 16, 2)], nmc=32768, steps=(1.0,))
    out = []
    for dyn in (0, 1):
        mci = build_mci(mcig, spec, nwalkers=1000, mode=0)
        mci.setDynamicScheduling(dyn)
        avg, err = mci.integrate(32768, False, False)
        wavg, werr = mci.walkerResults()
        out.append((avg, err, wavg.copy(), werr.copy(), mci.getAcceptanceRate(), mci.getX(walker=999)))
    for a, b in zip(out[0], out[1]):
        assert np.array_equal(np.asarray(a), np.asarray(b))


def test_forced_dynamic_scheduling_of_a_small_job_uses_the_whole_gpu(mcig):
    """setDynamicScheduling(1) on a job of fewer walker blocks than SMs once launched ONE persistent CTA (grid = floor(blocks / SMs) * SMs = 0 -> 1):
    correct but 70 times slower than the static launch (profiles/r02bd_ws_thin.log). Every block is resident now: within a factor of the static launch,
    and the same numbers."""
    res = []
    for dyn in (0, 1):
        mci = mcig.MCI(3)
        mci.setRngMode(0)
        mci.setSeed(11)
        mci.setNWalkers(8192)
        mci.addSamplingFunction(mcig.ThreeDimGaussianPDF())
        mci.addObservable(mcig.XSquared(), 0, 1)
        mci.setMRT2Step(1.0)
        mci.setDynamicScheduling(dyn)
        mci.integrate(20000, False, False)
        avg, err = mci.integrate(20000, False, False)
        res.append((avg[0], mci.timings()["walk_ms"]))
    assert res[0][0] == res[1][0]
    assert res[1][1] < 4.0*res[0][1], res


def test_file_dumps_match_reference_text(mcig, tmp_path):
    """storeObservablesOnFile / storeWalkerPositionsOnFile (src/MCIntegrator.cpp:495-542): the text the device path writes for walker
    0 in replay mode equals, character for character, what the reference wrote for the same run (tests/golden/dump_*.txt)."""
    import os
    spec = configs.RUNS["dump_files"]
    mci = build_mci(mcig, spec)
    op, wp = str(tmp_path / "observables.txt"), str(tmp_path / "walker.txt")
    mci.storeObservablesOnFile(op, configs.DUMP_OBS_FREQ)
    mci.storeWalkerPositionsOnFile(wp, configs.DUMP_WLK_FREQ)
    avg, err = mci.integrate(spec["nmc"], False, False)
    gold = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
    assert open(wp).read() == open(os.path.join(gold, "dump_walker.txt")).read()
    assert open(op).read() == open(os.path.join(gold, "dump_observables.txt")).read()
    # the dumps do not disturb the results, and can be switched off again
    mci2 = build_mci(mcig, spec)
    avg2, err2 = mci2.integrate(spec["nmc"], False, False)
    assert np.array_equal(avg, avg2) and np.array_equal(err, err2)
    mci.clearObservableFile()
    mci.clearWalkerFile()
    os.remove(op)
    mci.integrate(160, False, False)
    assert not os.path.exists(op)


@pytest.mark.parametrize("name", ["ndim_all64", "ndim_vec64_v4", "gauss_vec6_v3", "ms_sub16", "vec_ortho_types", "ndim_all96"])
def test_global_memory_placement_equals_shared_memory_in_philox_mode(name, mcig):
    """The global-memory state placement runs the same code over a different view: production-mode results must be bit-identical."""
    spec = configs.RUNS[name]
    out = []
    for placement in (1, 2):
        mci = build_mci(mcig, spec, nwalkers=96, mode=0, placement=placement)
        avg, err = mci.integrate(spec["nmc"], False, False)
        out.append((avg.copy(), err.copy(), mci.getAcceptanceRate(), list(mci.getX())))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]) and out[0][2:] == out[1][2:]


@pytest.mark.parametrize("mode", [0, 1])
@pytest.mark.parametrize("name", ["ndim_all64", "gauss_all", "srrd_exponential_all", "srrd_lognormal_all", "srrd_student_all"])
def test_streamed_draws_equal_register_draws(name, mode, mcig, monkeypatch):
    """All-moves over more than MCIG_STREAM_NDIM coordinates generate their Philox blocks on demand instead of holding every draw in
    registers. Same (group, walker, block) -> word mapping: lowering the threshold must not change a single bit, for any distribution."""
    spec = dict(configs.RUNS[name])
    out = []
    for defs in ("", "MCIG_STREAM_NDIM=1"):
        if defs:
            monkeypatch.setenv("MCIG_JIT_DEFINES", defs)
        else:
            monkeypatch.delenv("MCIG_JIT_DEFINES", raising=False)
        mci = build_mci(mcig, spec, nwalkers=64, mode=mode, placement=1)
        avg, err = mci.integrate(2048, False, False)
        out.append((avg.copy(), err.copy(), mci.getAcceptanceRate(), list(mci.getX())))
    assert np.array_equal(out[0][0], out[1][0]) and np.array_equal(out[0][1], out[1][1]) and out[0][2:] == out[1][2:]
    assert 0.02 < out[0][2] < 0.99


def test_thousand_dimensional_walkers_sample_the_right_distribution(mcig):
    """ndim = 1024 (the top of the reference's dimension sweeps, benchmark/bench_throughput_ndim_*): single-particle moves over
    global-memory walkers; <x_i^2> of the unit-variance-1/2 Gaussian is 0.5 for every coordinate (test/common/TestMCIFunctions.hpp)."""
    nd = 1024
    mci = mcig.MCI(nd)
    mci.setRngMode(0)
    mci.setNWalkers(2048)
    mci.setSeed(1337)
    mci.setTrialMove(mcig.SRRDType.Uniform, 1, 1, None)
    mci.setMRT2Step(0, 1.6)
    mci.setX([0.0]*nd)
    mci.addSamplingFunction(mcig.Gauss(nd))
    mci.addObservable(mcig.X2(nd), 0, 1)
    mci.integrate(20*nd, False, False)  # ~20 sweeps to equilibrate
    avg, _ = mci.integrate(40*nd, False, False)
    assert 0.3 < mci.getAcceptanceRate() < 0.7
    # 2048 walkers x 40 sweeps of correlated samples per coordinate: standard error ~ 0.5*sqrt(2)/sqrt(2048*40/4) ~ 0.005
    assert avg.shape == (nd,) and np.all(np.abs(avg - 0.5) < 0.03), (avg.min(), avg.max())
    assert abs(avg.mean() - 0.5) < 2e-3


@pytest.mark.parametrize("name", configs.CALLBACK_RUNS)
def test_step_callback_matches_reference(name, mcig, golden_callback):
    """MCI::setCallback as a device functor: called once at the start of every sampling run and after every move (calibration and
    decorrelation runs included), with the pre-commit state. The four sums the reference-side harness callback folds must agree."""
    spec = configs.RUNS[name]
    g = golden_callback[name]
    mci = build_mci(mcig, spec)
    mci.setCallback(mcig.StepCallback("FoldCallback"), 6)
    avg, err = mci.integrate(spec["nmc"], spec.get("do_find", False), spec.get("do_decorr", False))
    buf = mci.getCallbackBuffer()
    ref = fromhex(g["buf"])
    assert buf[0] == ref[0] and buf[1] == ref[1], (buf, ref)
    assert _close(buf[2:4], ref[2:4], 1e-12, atol=1e-12), (buf, ref)
    assert mci.getAcceptanceRate() == float.fromhex(g["acc_rate"])
    assert buf[5] == spec["nmc"] - 1  # the last call saw the last step of the main run
    if not spec.get("do_find", False) and not spec.get("do_decorr", False):
        n = spec["nmc"]
        assert buf[4] == n*(n - 1)/2 - 1  # steps -1 (initializeSampling), 0 .. n-1
    mci.clearCallback()
    avg2, _ = mci.integrate(spec["nmc"], False, False)
    assert avg2.shape == avg.shape


def test_step_callback_many_walkers_philox(mcig):
    nd, W, n = 3, 4096, 1000
    mci = mcig.MCI(nd)
    mci.setRngMode(0)
    mci.setNWalkers(W)
    mci.setSeed(5)
    mci.addSamplingFunction(mcig.ThreeDimGaussianPDF())
    mci.addObservable(mcig.XSquared(), 0, 1)
    mci.setMRT2Step(1.0)
    from prod import register_test_plugins
    register_test_plugins(mcig)
    mci.setCallback(mcig.StepCallback("FoldCallback"), 6*W)
    mci.integrate(n, False, False)
    buf = mci.getCallbackBuffer().reshape(W, 6)
    assert np.all(buf[:, 0] == n + 1) and np.all(buf[:, 5] == n - 1)
    assert buf[:, 1].sum() - W == round(mci.getAcceptanceRate()*W*n)  # per-walker accept counts add up to the global rate


@pytest.mark.parametrize("name", ["vec_exp4", "ndim_vec16", "ndim_vec64_v4", "ndim_vec300_v3_types"])
def test_lazy_accumulation_against_the_reference(name, mcig, golden_runs):
    """Value x dwell-time accumulation forced on in replay mode: same trajectory as the reference (bit-exact), sums equal to the
    reference's step-by-step additions up to the rounding of the summation order (1e-12 of the observable's scale)."""
    spec = configs.RUNS[name]
    g = golden_runs[name]
    mci = build_mci(mcig, spec)
    mci.setLazyAccumulation(2)
    avg, err = mci.integrate(spec["nmc"], spec.get("do_find", False), spec.get("do_decorr", False))
    assert "LazyAccu" in mci.kernelSource()
    assert mci.getAcceptanceRate() == float.fromhex(g["acc_rate"]) and list(mci.getX()) == fromhex(g["x_final"])
    ref = np.array(fromhex(g["avg"]))
    scale = max(1.0, float(np.max(np.abs(ref))))
    assert np.max(np.abs(avg - ref)) <= 1e-12*scale, np.max(np.abs(avg - ref))
    assert _close(err, fromhex(g["err"]), 1e-9, atol=1e-15)


@pytest.mark.parametrize("name", ["block16", "full_uncorr", "mixed", "block_skip_big", "tp3g_small", "gauss_all", "srrd_gamma_vec", "full_two_samples"])
@pytest.mark.parametrize("dyn", [0, 1])
def test_fused_one_pass_estimator_is_the_reference_bit_for_bit(name, dyn, mcig, oracle, golden_runs):
    """Block / Full accumulators with the uncorrelated estimator never touch HBM: the walker keeps sum x and sum x^2 of what it would have
    stored (src/Estimators.cpp:36-56, 125-155: one pass, left to right) in registers. In replay mode averages AND errors are then the
    reference's to the last bit where the estimator is the 1-D one (std::accumulate / inner_product order), and within the usual
    tolerances otherwise; keeping the series as well (setKeepSamples) must not change a bit, nor may the chunked (dynamic) schedule."""
    spec = configs.RUNS[name]
    g = golden_runs[name]
    if name in ("srrd_gamma_vec",) and dyn:
        pytest.skip("state-memory walkers have no dynamically scheduled kernel")
    res = []
    for keep in (False, True):
        mci = build_mci(mcig, spec)
        mci.setKeepSamples(keep)
        if dyn:
            mci.setStatePlacement(0)
            mci.setDynamicScheduling(1)
        avg, err = mci.integrate(spec["nmc"], False, False)
        src = mci.kernelSource()
        assert ("Accu<" in src) and ((",2> a" in src or ",2,false> a" in src) != keep)
        res.append((avg.copy(), err.copy()))
        if not keep:
            from mcintegratorplusplus_b200._capi import McigError
            fused = [i for i, o in enumerate(spec["obs"]) if tuple(o)[1] >= 1 and (tuple(o)[4] if len(tuple(o)) > 4 else orc.default_estim(tuple(o)[1])) == orc.EST_UNCORRELATED]
            assert fused
            with pytest.raises(McigError, match="were not stored"):
                mci.obsData(fused[0], walker=0, nobs=1)
    assert np.array_equal(res[0][0], res[1][0]) and np.array_equal(res[0][1], res[1][1])
    assert _close(res[0][0], fromhex(g["avg"]), AVG_RTOL) and _close(res[0][1], fromhex(g["err"]), ERR_RTOL, atol=1e-18)
