"""Named run configurations shared by oracle/gen_golden.py (which runs them through the UNMODIFIED reference and commits
the outputs under tests/golden/) and by the parity tests. Each maps to keyword arguments of oracle/orc.py:make_config.

Sources of the workloads (paths relative to the reference repository):
  c1_*      BASELINE.json configs[0] pinned as SURVEY.md §8(d) C1 (3-D Gaussian / x^2, Simple, all-move)
  mixed     benchmark/bench_integrate_mixed/main.cpp:28-46
  tp3g      benchmark/bench_throughput_3G/main.cpp:23-38 (1-D Exp / X1D, Block(20), step 3.185) at a small Nmc
  ut4_*     test/ut4/main.cpp:33-73   ut5_*  test/ut5/main.cpp:41-137   ut2_*  test/ut2/main.cpp:56-91
  ndim_vec  benchmark/bench_throughput_ndim_single/main.cpp:26-50
"""
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "oracle"))
import orc  # noqa: E402


def _alt(nd):
    return [0.1 if j % 2 == 0 else -0.05 for j in range(nd)]


RUNS = {
    # --- all-move, 3-D Gaussian (the north-star integrand)
    "c1_simple": dict(ndim=3, seed=5649871, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 0, 1)], nmc=100000, steps=(1.0,)),
    "c1_simple_short": dict(ndim=3, seed=1337, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 0, 1)], nmc=4096, steps=(1.0,)),
    "full_mj": dict(ndim=3, seed=1337, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 1, 1), (orc.OBS_XYZSQUARED, 1, 1)], nmc=65536, steps=(1.0,)),
    "full_fc": dict(ndim=3, seed=1337, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 1, 1), (orc.OBS_XYZSQUARED, 1, 1)], nmc=60000, steps=(1.0,)),
    "full_uncorr": dict(ndim=3, seed=1337, pdf_id=orc.PDF_GAUSS3D,
                        obs=[(orc.OBS_XSQUARED, 1, 1, True, orc.EST_UNCORRELATED), (orc.OBS_XYZSQUARED, 1, 1, True, orc.EST_UNCORRELATED)],
                        nmc=65536, steps=(1.0,)),
    "block16": dict(ndim=3, seed=1337, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 16, 1), (orc.OBS_XYZSQUARED, 16, 1)], nmc=65536, steps=(1.0,)),
    "block8_skip2_corr": dict(ndim=3, seed=1337, pdf_id=orc.PDF_GAUSS3D,
                              obs=[(orc.OBS_XSQUARED, 8, 2, True, orc.EST_CORRELATED), (orc.OBS_XYZSQUARED, 8, 2, True, orc.EST_CORRELATED)],
                              nmc=65536, steps=(1.0,)),
    "mixed": dict(ndim=3, seed=5649871, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XND, 0, 1), (orc.OBS_XSQUARED, 1, 5), (orc.OBS_XYZSQUARED, 5, 2)],
                  nmc=100000, steps=(1.0,)),
    "tp3g_small": dict(ndim=1, seed=1337, pdf_id=orc.PDF_EXP1D, obs=[(orc.OBS_X1D, 20, 1)], nmc=100000, steps=(3.185,)),
    # --- automatic calibration + decorrelation
    "auto_default": dict(ndim=3, seed=1337, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 1, 1)], nmc=65536, x0=(5., -5., 10.), do_find=True, do_decorr=True),
    "ut2_irange": dict(ndim=3, seed=5649871, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 1, 1)], nmc=10000, x0=(5., -5., 10.), lb=-5., ub=5.,
                       do_find=True, do_decorr=True),
    "ut4_fixed": dict(ndim=3, seed=1337, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 1, 1), (orc.OBS_XYZSQUARED, 16, 1)], nmc=16384,
                      x0=(5., -5., 10.), nfind=20, ndecorr=2000, do_find=True, do_decorr=True),
    "ut3_two_obs": dict(ndim=3, seed=5649871, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 1, 1), (orc.OBS_XYZSQUARED, 1, 1)], nmc=10000,
                        x0=(5., -5., 10.), do_find=True, do_decorr=True),
    # --- single-vector moves / typed step sizes / selective updates
    "vec_exp4": dict(ndim=4, seed=1337, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 0, 1)], nmc=50000, move_type=orc.MOVE_VEC, veclen=1, steps=(3.0,), x0=_alt(4)),
    "vec_gauss3_auto": dict(ndim=3, seed=1337, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_XSQUARED, 1, 3), (orc.OBS_X2, 1, 3)], nmc=32768*3, move_type=orc.MOVE_VEC,
                            veclen=1, do_find=True, do_decorr=True),
    "vec3_types": dict(ndim=6, seed=1337, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2, 1, 2), (orc.OBS_UPDXND, 4, 1)], nmc=32768, move_type=orc.MOVE_VEC, veclen=3,
                       ntypes=2, type_ends=[3, 6], steps=(0.4, 0.8), do_find=True, do_decorr=True),
    "all_types": dict(ndim=4, seed=77, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2SUM, 1, 1), (orc.OBS_POLYNOM, 0, 1)], nmc=16384, ntypes=2, type_ends=[1, 4],
                      steps=(0.9, 0.5)),
    "ndim_vec16": dict(ndim=16, seed=1337, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 0, 1)], nmc=20000, move_type=orc.MOVE_VEC, veclen=1, steps=(3.0,),
                       x0=_alt(16), ndecorr=2000, do_decorr=True),
    "ndim_all16": dict(ndim=16, seed=1337, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 0, 1)], nmc=20000, steps=(0.4,), x0=_alt(16)),
    "ndim_all64": dict(ndim=64, seed=4242, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2SUM, 8, 1)], nmc=8192, steps=(0.12,), x0=_alt(64)),
    "ndim_vec64_v4": dict(ndim=64, seed=4242, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2SUM, 1, 4), (orc.OBS_XND, 0, 1)], nmc=16384, move_type=orc.MOVE_VEC,
                          veclen=4, steps=(0.6,), x0=_alt(64)),
    # --- walkers beyond the register / shared-memory budget (the reference's dimension sweeps go to 1024): streamed all-move
    #     draws above 64 coordinates, global-memory state placement once a warp of walkers exceeds 227 KiB of shared memory
    "ndim_all96": dict(ndim=96, seed=11, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2SUM, 8, 1), (orc.OBS_XND, 0, 1)], nmc=2048, steps=(0.1,), x0=_alt(96)),
    "ndim_gauss_all96": dict(ndim=96, seed=12, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2SUM, 1, 1)], nmc=2048, srrd=orc.SRRD_GAUSSIAN, steps=(0.06,), x0=_alt(96)),
    "ndim_vec256": dict(ndim=256, seed=13, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2SUM, 1, 1), (orc.OBS_UPDXND, 8, 2)], nmc=4096, move_type=orc.MOVE_VEC, veclen=1,
                        steps=(1.2,), x0=_alt(256)),
    "ndim_vec300_v3_types": dict(ndim=300, seed=14, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 0, 1)], nmc=4096, move_type=orc.MOVE_VEC, veclen=3, ntypes=2,
                                 type_ends=[150, 300], steps=(2.0, 1.0), x0=_alt(300), lb=-8., ub=8., nfind=-20, ndecorr=-2000, do_find=True, do_decorr=True),
    "ndim_all300": dict(ndim=300, seed=15, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 16, 1), (orc.OBS_X2SUM, 1, 1)], nmc=1024, steps=(0.05,), x0=_alt(300)),
    "ndim_ms256": dict(ndim=256, seed=16, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 20, 1)], nmc=400, move_type=orc.MOVE_MULTISTEP, veclen=1, steps=(0.2,),
                       x0=_alt(256)),
    "ndim_all1024": dict(ndim=1024, seed=17, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_XND, 0, 1), (orc.OBS_X2SUM, 4, 1)], nmc=256, steps=(0.03,), x0=_alt(1024)),
    "ndim_vec1024": dict(ndim=1024, seed=18, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2, 1, 8)], nmc=4096, move_type=orc.MOVE_VEC, veclen=2, steps=(1.0,), x0=_alt(1024)),
    # --- dependent observables (include/mci/DependentObservableInterface.hpp; reference side: oracle/ref_harness.cpp HarnessDepObs)
    "dep_obs_all": dict(ndim=3, seed=301, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_XSQUARED, 1, 1), (orc.OBS_DEPENDENT, 4, 2)], nmc=8192, steps=(0.8,)),
    "dep_obs_vec": dict(ndim=4, seed=302, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2SUM, 0, 1), (orc.OBS_DEPENDENT, 1, 1), (orc.OBS_XND, 0, 1)], nmc=8192,
                        move_type=orc.MOVE_VEC, veclen=1, steps=(1.1,)),
    "dep_obs_auto": dict(ndim=3, seed=303, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2SUM, 1, 1), (orc.OBS_DEPENDENT, 1, 1)], nmc=4096, x0=(1., -1., 2.),
                         do_find=True, do_decorr=True),
    # --- MultiStepMove
    "ms_default4": dict(ndim=4, seed=1337, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_XND, 0, 1)], nmc=50000, move_type=orc.MOVE_MULTISTEP, veclen=1, steps=(0.7,)),
    "ms_sub_ut5": dict(ndim=3, seed=1337, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_XSQUARED, 1, 1), (orc.OBS_X2, 1, 3)], nmc=32768*3, move_type=orc.MOVE_MULTISTEP,
                       veclen=1, ms_nsteps=3, ms_sub_pdf_id=orc.PDF_EXPND, steps=(0.1,), do_find=True, do_decorr=True, target_acc=0.85),
    "ms_sub16": dict(ndim=16, seed=99, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_XND, 20, 1)], nmc=4000, move_type=orc.MOVE_MULTISTEP, veclen=1,
                     ms_sub_pdf_id=orc.PDF_EXPND, steps=(0.5,), x0=_alt(16)),
    "ms_nosub8": dict(ndim=8, seed=5, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 20, 1)], nmc=4000, move_type=orc.MOVE_MULTISTEP, veclen=1, steps=(0.3,),
                      x0=_alt(8)),
    # --- the shapes BASELINE configs[2] (C3) actually runs: ExpNDPDF / Gauss + XND with Block(20) at ndim 32 and 64
    #     (benchmark/bench_throughput_ndim_single/main.cpp:26-50, src/MultiStepMove.cpp:6-47); every state placement is tested
    "ms_sub32_b20": dict(ndim=32, seed=3201, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_XND, 20, 1)], nmc=2000, move_type=orc.MOVE_MULTISTEP, veclen=1,
                         ms_sub_pdf_id=orc.PDF_EXPND, steps=(0.5,), x0=_alt(32)),
    "ms_sub64_b20": dict(ndim=64, seed=6401, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_XND, 20, 1)], nmc=1000, move_type=orc.MOVE_MULTISTEP, veclen=1,
                         ms_sub_pdf_id=orc.PDF_EXPND, steps=(0.5,), x0=_alt(64)),
    "ms_nosub32_b20": dict(ndim=32, seed=3202, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 20, 1)], nmc=2000, move_type=orc.MOVE_MULTISTEP, veclen=1,
                           steps=(0.05,), x0=_alt(32)),
    "ms_nosub64_b20": dict(ndim=64, seed=6402, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 20, 1)], nmc=1000, move_type=orc.MOVE_MULTISTEP, veclen=1,
                           steps=(0.03,), x0=_alt(64)),
    "ndim_all32_b20": dict(ndim=32, seed=3203, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 20, 1)], nmc=20000, steps=(0.53,), x0=_alt(32)),
    "ndim_all64_b20": dict(ndim=64, seed=6403, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 20, 1)], nmc=10000, steps=(0.375,), x0=_alt(64)),
    "ndim_gauss_all32": dict(ndim=32, seed=3204, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2SUM, 1, 1), (orc.OBS_XND, 0, 1)], nmc=16384, steps=(0.3,), x0=_alt(32)),
    "ndim_vec32_b20": dict(ndim=32, seed=3205, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 20, 1)], nmc=20000, move_type=orc.MOVE_VEC, veclen=1, steps=(3.0,),
                           x0=_alt(32)),
    "ndim_vec64_b20": dict(ndim=64, seed=6405, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 20, 1)], nmc=20000, move_type=orc.MOVE_VEC, veclen=1, steps=(3.0,),
                           x0=_alt(64)),
    # --- lane-split walkers (state placement 3, automatic for eligible all-moves from ndim 64 on): typed steps, periodic domain, Simple / Full /
    #     Block accumulators of element-wise observables, automatic calibration + decorrelation
    "lanes_all128_auto": dict(ndim=128, seed=12801, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_XND, 0, 1), (orc.OBS_X2, 1, 2), (orc.OBS_UPDXND, 8, 1)], nmc=4096,
                              ntypes=2, type_ends=[64, 128], steps=(0.2, 0.1), lb=-3., ub=3., x0=_alt(128), nfind=-20, ndecorr=-2000, do_find=True, do_decorr=True),
    "lanes_all8_b4": dict(ndim=8, seed=801, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_XND, 4, 1), (orc.OBS_X2, 1, 1, False, orc.EST_UNCORRELATED)], nmc=2048, steps=(0.6,),
                          x0=_alt(8)),  # observables of <= 8 components with the uncorrelated estimator: the ones the walk would fuse; lane slices store their series
    "lanes_all192": dict(ndim=192, seed=19201, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_X2, 16, 1, True, orc.EST_CORRELATED)], nmc=2048, steps=(0.15,), x0=_alt(192)),
    # --- Gaussian proposals (SRRDType::Gaussian: std::normal_distribution, polar method with a cached value)
    "gauss_all": dict(ndim=3, seed=99, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 16, 1)], nmc=16384, srrd=orc.SRRD_GAUSSIAN, steps=(0.6,)),
    "gauss_all_auto": dict(ndim=3, seed=98, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 1, 1)], nmc=8192, srrd=orc.SRRD_GAUSSIAN, x0=(2., -2., 1.),
                           do_find=True, do_decorr=True),
    "gauss_vec5": dict(ndim=5, seed=97, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2, 1, 1), (orc.OBS_XND, 0, 1)], nmc=20000, move_type=orc.MOVE_VEC, veclen=1,
                       srrd=orc.SRRD_GAUSSIAN, steps=(1.2,)),
    "gauss_vec6_v3": dict(ndim=6, seed=96, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_X2SUM, 8, 1)], nmc=16384, move_type=orc.MOVE_VEC, veclen=3,
                          srrd=orc.SRRD_GAUSSIAN, steps=(0.9,), lb=-4., ub=4.),
    # --- the other eight SRRD proposal distributions (createSymRRD defaults). Reference values come from the harness only (goldens):
    #     the C oracle restates uniform and normal; the CUDA path replays the libstdc++ outputs of all of them.
    "srrd_student_all": dict(ndim=3, seed=201, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 1, 1)], nmc=8192, srrd=2, steps=(0.4,)),
    "srrd_cauchy_vec": dict(ndim=4, seed=202, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2, 8, 1)], nmc=8192, move_type=orc.MOVE_VEC, veclen=2, srrd=3, steps=(0.5,)),
    "srrd_exponential_all": dict(ndim=3, seed=203, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XYZSQUARED, 0, 1)], nmc=8192, srrd=4, steps=(0.6,), do_find=True, do_decorr=True),
    "srrd_gamma_vec": dict(ndim=3, seed=204, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 16, 1)], nmc=8192, move_type=orc.MOVE_VEC, veclen=1, srrd=5, steps=(1.0,)),
    "srrd_weibull_all": dict(ndim=2, seed=205, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2SUM, 1, 1)], nmc=8192, srrd=6, steps=(0.7,)),
    "srrd_lognormal_all": dict(ndim=3, seed=206, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 1, 1)], nmc=8192, srrd=7, steps=(0.3,)),
    "srrd_chisq_vec": dict(ndim=6, seed=207, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2, 1, 2)], nmc=8192, move_type=orc.MOVE_VEC, veclen=3, srrd=8, steps=(0.4,)),
    "srrd_fisher_all": dict(ndim=1, seed=208, pdf_id=orc.PDF_EXP1D, obs=[(orc.OBS_X1D, 32, 1)], nmc=8192, srrd=9, steps=(0.5,), lb=-6., ub=6.),
    # --- moves built around a pre-made distribution with non-default parameters (include/mci/SRRDAllMove.hpp:45-58; ut5's StudentAllMove(ndim, 0.05,
    #     &student_t(2)) with its automatic calibration, test/ut5/main.cpp:110-125). srrd_par: Gaussian stddev; Student n; Cauchy b; Exponential lambda;
    #     Gamma alpha, beta; Weibull a, b; Lognormal m, s; Chisq n; Fisher m, n. The *_vec runs pin a quirk: the reference's vec-move clone drops the
    #     distribution (include/mci/SRRDVecMove.hpp:30-33), so MCI samples those with the DEFAULT parameters whatever was passed.
    "ut5_student2_all": dict(ndim=3, seed=5005, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_XSQUARED, 1, 1), (orc.OBS_XND, 16, 1)], nmc=8192, srrd=2, srrd_par=(2.0,),
                             x0=(0.5, -0.5, 0.25), do_find=True, do_decorr=True),
    "par_gauss_all": dict(ndim=3, seed=301, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 1, 1)], nmc=4096, srrd=1, srrd_par=(0.5,), steps=(1.3,)),
    "par_student5_vec": dict(ndim=4, seed=302, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2, 8, 1)], nmc=4096, move_type=orc.MOVE_VEC, veclen=2, srrd=2, srrd_par=(5.0,),
                             steps=(0.6,)),
    "par_cauchy_all": dict(ndim=2, seed=303, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2SUM, 1, 1)], nmc=4096, srrd=3, srrd_par=(0.25,), steps=(1.0,)),
    "par_exponential_all": dict(ndim=3, seed=304, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XYZSQUARED, 0, 1)], nmc=4096, srrd=4, srrd_par=(2.5,), steps=(1.5,)),
    "par_gamma_vec": dict(ndim=3, seed=305, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 16, 1)], nmc=4096, move_type=orc.MOVE_VEC, veclen=1, srrd=5,
                          srrd_par=(2.5, 0.5), steps=(1.0,)),
    "par_weibull_all": dict(ndim=2, seed=306, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2SUM, 1, 1)], nmc=4096, srrd=6, srrd_par=(1.5, 0.8), steps=(0.7,)),
    "par_lognormal_all": dict(ndim=3, seed=307, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 1, 1)], nmc=4096, srrd=7, srrd_par=(-0.5, 0.6), steps=(0.5,)),
    "par_chisq_vec": dict(ndim=6, seed=308, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2, 1, 2)], nmc=4096, move_type=orc.MOVE_VEC, veclen=3, srrd=8, srrd_par=(3.0,),
                          steps=(0.2,)),
    "par_fisher_all": dict(ndim=1, seed=309, pdf_id=orc.PDF_EXP1D, obs=[(orc.OBS_X1D, 32, 1)], nmc=4096, srrd=9, srrd_par=(4.0, 6.0), steps=(0.5,), lb=-6., ub=6.),
    # --- a user-defined trial move (TrialMoveInterface subclass in the reference harness, device functor here): drifted uniform proposal whose
    #     acceptance factor is 0 or 1 (oracle/ref_harness.cpp: HarnessDriftMove; tests/prod.py: DRIFT_MOVE_SRC); srrd_par = (drift,)
    "user_drift_all": dict(ndim=3, seed=7101, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 1, 1), (orc.OBS_XYZSQUARED, 8, 1)], nmc=8192,
                           move_type=orc.MOVE_USER_DRIFT, srrd_par=(0.15,), steps=(0.9,), x0=(0.3, -0.2, 0.1)),
    "user_drift_types_ortho": dict(ndim=4, seed=7102, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2, 1, 1)], nmc=4096, move_type=orc.MOVE_USER_DRIFT, srrd_par=(-0.05,),
                                   ntypes=2, type_ends=[1, 4], steps=(1.1, 0.7), lb=-2.5, ub=2.5, x0=(1., -1., 0.5, 0.), do_decorr=True),
    "user_drift_all24": dict(ndim=24, seed=7103, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 16, 1)], nmc=2048, move_type=orc.MOVE_USER_DRIFT, srrd_par=(0.01,),
                             steps=(0.2,), x0=_alt(24)),
    # --- edge cases
    "vec_ortho_types": dict(ndim=6, seed=31, pdf_id=orc.PDF_EXPND, obs=[(orc.OBS_XND, 4, 1), (orc.OBS_X2SUM, 1, 1)], nmc=8192, move_type=orc.MOVE_VEC, veclen=2,
                            ntypes=3, type_ends=[2, 4, 6], steps=(1.5, 2.5, 3.5), lb=[-2., -3., -2., -3., -2., -3.], ub=[2., 3., 2., 3., 2., 3.],
                            x0=[1.9, -2.9, 0.1, 0.2, -1.9, 2.9]),
    "ms_ortho": dict(ndim=4, seed=8, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2, 2, 1)], nmc=4096, move_type=orc.MOVE_MULTISTEP, veclen=1, ms_nsteps=6,
                     ms_sub_pdf_id=orc.PDF_GAUSS, steps=(0.8,), lb=-1.5, ub=1.5, x0=[1.4, -1.4, 0.3, 0.]),
    "no_obs": dict(ndim=3, seed=77, pdf_id=orc.PDF_GAUSS3D, obs=[], nmc=5000, steps=(0.7,), x0=(1., 2., 3.)),
    "calib_only": dict(ndim=3, seed=78, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 1, 1)], nmc=0, x0=(1., 2., 3.), do_find=True, do_decorr=True),
    "nmc_one": dict(ndim=3, seed=79, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 0, 1), (orc.OBS_XYZSQUARED, 1, 1, False, orc.EST_NOOP)], nmc=1, steps=(1.0,),
                    x0=(0.5, 0.5, 0.5)),
    "nskip_gt_nmc": dict(ndim=3, seed=80, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 0, 1000), (orc.OBS_XYZSQUARED, 1, 7, False, orc.EST_NOOP)], nmc=500,
                         steps=(1.0,), x0=(0.5, 0.5, 0.5)),
    "full_two_samples": dict(ndim=1, seed=81, pdf_id=orc.PDF_EXP1D, obs=[(orc.OBS_X1D, 1, 1, False, orc.EST_UNCORRELATED)], nmc=2, steps=(2.0,)),
    "vec_full_len": dict(ndim=3, seed=82, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_XSQUARED, 1, 1)], nmc=4096, move_type=orc.MOVE_VEC, veclen=3, steps=(0.9,)),
    "block_skip_big": dict(ndim=2, seed=83, pdf_id=orc.PDF_GAUSS, obs=[(orc.OBS_X2, 64, 3, True, orc.EST_UNCORRELATED)], nmc=64*3*40, steps=(1.1,)),
    # --- periodic text dumps (test/main.cpp:108-113 uses both with freq 100)
    "dump_files": dict(ndim=3, seed=909, pdf_id=orc.PDF_GAUSS3D, obs=[(orc.OBS_XSQUARED, 1, 1), (orc.OBS_XYZSQUARED, 4, 2), (orc.OBS_CONSTVAL, 0, 5)],
                       nmc=4000, steps=(0.9,), x0=(0.2, -0.3, 0.4)),
    # --- no sampling function (plain MC over a box), ex_basic
    "nopdf_box": dict(ndim=3, seed=42, pdf_id=orc.PDF_NONE, obs=[(orc.OBS_GAUSSXSQUARED, 1, 1)], nmc=16384, lb=-5., ub=5.),
    "exbasic_1": dict(ndim=1, seed=7, pdf_id=orc.PDF_NONE, obs=[(orc.OBS_PARABOLA, 1, 1)], nmc=100000, lb=-1., ub=3., x0=(-0.5,), steps=(0.25,), target_acc=0.7,
                      do_find=True, do_decorr=True),
    "exbasic_2": dict(ndim=1, seed=7, pdf_id=orc.PDF_NORMLINE, obs=[(orc.OBS_NORMPARABOLA, 1, 1)], nmc=100000, lb=-1., ub=3., x0=(-0.5,), steps=(0.25,),
                      target_acc=0.7, do_find=True, do_decorr=True),
}


DUMP_OBS_FREQ, DUMP_WLK_FREQ = 100, 250


def in_oracle(name):
    """The plain-C oracle restates the uniform and normal proposal distributions; the rest is pinned by the goldens only."""
    return RUNS[name].get("srrd", 0) < 2 and not RUNS[name].get("srrd_par") and RUNS[name].get("move_type", 0) != orc.MOVE_USER_DRIFT and all(tuple(o)[0] != orc.OBS_DEPENDENT for o in RUNS[name]["obs"])


# configurations whose step callback sums are pinned by the reference (oracle/ref_harness.cpp: mciref_run_callback)
# the C3 shapes whose kernels are footprint- / register-limited: replayed on every state placement (tests/test_walk_parity.py)
C3_SHAPES = ["ms_sub32_b20", "ms_sub64_b20", "ms_nosub32_b20", "ms_nosub64_b20", "ndim_all32_b20", "ndim_all64_b20", "ndim_gauss_all32",
             "ndim_vec32_b20", "ndim_vec64_b20"]

# configurations eligible for lane-split walkers, with the lane count placement 3 gives them
LANE_SPLIT = {"lanes_all8_b4": 2, "ndim_all16": 2, "ndim_all32_b20": 2, "ndim_all64_b20": 4, "lanes_all128_auto": 8, "lanes_all192": 8}

CALLBACK_RUNS = ["c1_simple_short", "vec_exp4", "ms_sub16", "auto_default", "ut2_irange", "nopdf_box", "gauss_vec6_v3"]


def make(name):
    kw = dict(RUNS[name])
    ndim, seed, pdf_id, obs, nmc = kw.pop("ndim"), kw.pop("seed"), kw.pop("pdf_id"), kw.pop("obs"), kw.pop("nmc")
    return orc.make_config(ndim, seed, pdf_id, obs, nmc, **kw)


# TestWalk data sets for estimator parity: (pdf 0 SLATER / 1 GAUSS, nmc, ndim, step, change_prob, srand seed)
# bench_estimators (benchmark/bench_estimators/main.cpp:31-51) at reduced size, ut1 (test/ut1/main.cpp:228-230)
WALKS = {
    "ut1_gauss": (1, 32768, 2, 2.0, 0.5, 1337),
    "est_1d": (0, 65536, 1, 1.59, 1.0, 1337),
    "est_16d": (0, 4096, 16, 0.313, 1.0, 1337),
    "est_1d_npow2": (0, 60000, 1, 1.59, 1.0, 4711),
    "est_3d_npow2": (1, 50001, 3, 0.8, 1.0, 99),
    "est_small": (0, 346, 3, 0.8, 1.0, 5),
}
