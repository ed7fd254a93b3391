"""CPU: the C-ABI library loads, exports every symbol include/mcig.h declares, generates and JIT-compiles (NVRTC needs no
GPU) the walk kernel for the headline configurations, and fails LOUDLY — no CPU fallback — when asked to compute."""
import os
import re

import pytest

import configs
from prod import build_mci

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_header_symbols_exported(mcig):
    hdr = open(os.path.join(ROOT, "include", "mcig.h")).read()
    declared = set(re.findall(r"\b(mcig_[a-z0-9_]+)\s*\(", hdr)) - {"mcig_allreduce_fn"}
    from mcintegratorplusplus_b200 import _capi
    lib = _capi.lib()
    assert declared, "no declarations parsed"
    for name in sorted(declared):
        assert hasattr(lib, name), "libmcig.so does not export %s" % name
    assert declared == set(_capi.SIGNATURES), "ctypes table out of sync with include/mcig.h"
    assert lib.mcig_version() == 100


def test_builtin_plugins_registered(mcig):
    from mcintegratorplusplus_b200 import _capi
    lib = _capi.lib()
    for n in ("ThreeDimGaussianPDF", "Gauss", "Exp1DPDF", "ExpNDPDF", "NormalizedLine"):
        assert lib.mcig_lookup_plugin(0, n.encode()) >= 0
    for n in ("XSquared", "GaussXSquared", "XYZSquared", "X1D", "XND", "UpdateableXND", "Constval", "Polynom", "X2Sum", "X2", "Parabola",
              "NormalizedParabola"):
        assert lib.mcig_lookup_plugin(1, n.encode()) >= 0
    assert lib.mcig_lookup_plugin(0, b"nope") == -1


@pytest.mark.parametrize("name", ["c1_simple", "mixed", "vec_exp4", "ndim_vec16", "ms_sub_ut5", "ms_sub16", "nopdf_box", "ut2_irange", "vec3_types",
                                  "ndim_all64", "ndim_vec64_v4", "gauss_all", "gauss_vec5", "gauss_vec6_v3", "srrd_student_all", "srrd_cauchy_vec", "srrd_exponential_all", "srrd_lognormal_all",
                                  "srrd_chisq_vec", "srrd_fisher_all"])
@pytest.mark.parametrize("mode", [0, 2])
def test_jit_compiles_without_gpu(name, mode, mcig, tmp_path, monkeypatch):
    monkeypatch.setenv("MCIG_CACHE_DIR", str(tmp_path))
    mci = build_mci(mcig, configs.RUNS[name], mode=mode)
    src = mci.kernelSource()
    assert "mcig_walk" in src and "struct Glue" in src
    mci.prebuild()
    assert any(f.endswith(".cubin") for f in os.listdir(tmp_path))


def test_user_plugin_source_compiles(mcig, tmp_path, monkeypatch):
    monkeypatch.setenv("MCIG_CACHE_DIR", str(tmp_path))
    src = """
struct MyQuartic { // exp(-(a*sum x^4)), a = par[0]
    static constexpr int NPAR = 1; static constexpr bool HAS_UPDATE = false; static constexpr bool ELEMENTWISE = false;
    const double * par;
    template <class X, class P> __device__ void protoFunction(const X & in, P & pv) const { double s = 0.; for (int i = 0; i < 2; ++i) { const double q = in[i]*in[i]; s += q*q; } pv[0] = par[0]*s; }
    template <class P> __device__ double samplingFunction(const P & pv) const { return mcig::exp(-pv[0]); }
    template <class PO, class PN> __device__ double acceptanceFunction(const PO & po, const PN & pn) const { return mcig::exp(po[0] - pn[0]); }
};
"""
    mcig.register_plugin(0, "MyQuartic", "MyQuartic", src, ndim=2, nvalues=1, npar=1)
    mci = mcig.MCI(2)
    mci.addSamplingFunction(mcig.SamplingFunction("MyQuartic", par=[0.5]))
    mci.addObservable(mcig.X2Sum(2), 16, 1)
    mci.prebuild()
    assert "MyQuartic" in mci.kernelSource()


def test_argument_errors_mirror_reference(mcig):
    from mcintegratorplusplus_b200._capi import McigError
    mci = mcig.MCI(3)
    with pytest.raises(McigError, match="number of inputs is not equal"):
        mci.addSamplingFunction(mcig.Exp1DPDF())  # 1-D pdf into a 3-D MCI: src/MCIntegrator.cpp:486-488
    with pytest.raises(McigError, match="number of inputs is not equal"):
        mci.addObservable(mcig.X1D())  # src/MCIntegrator.cpp:456-458
    with pytest.raises(McigError, match="requires estimator with error calculation"):
        mci.addObservable(mcig.XSquared(), 1, 1, True, mcig.EstimatorType.Noop)  # src/MCIntegrator.cpp:459-461
    with pytest.raises(McigError, match="multiple of passed veclen"):
        mci.setTrialMove(mcig.SRRDType.Uniform, 2)  # src/MCIntegrator.cpp:435-437
    with pytest.raises(McigError, match="truly greater"):
        mci.setIRange(1.0, -1.0)  # src/OrthoPeriodicDomain.cpp:6-13
    with pytest.raises(McigError, match="infinite domain"):
        mci.addObservable(mcig.XSquared())
        mci.integrate(100)  # no pdf, unbound domain: src/MCIntegrator.cpp:45-47


def test_no_cpu_fallback(mcig):
    """Without a GPU every compute entry point must raise, never silently compute on the host."""
    from mcintegratorplusplus_b200 import _capi
    if _capi.lib().mcig_device_count() > 0:
        pytest.skip("a GPU is present")
    mci = build_mci(mcig, configs.RUNS["c1_simple_short"], mode=0)
    with pytest.raises(_capi.McigError, match="no CUDA device"):
        mci.integrate(128, False, False)
    import numpy as np
    with pytest.raises(_capi.McigError, match="no CUDA device"):
        mcig.estimate(mcig.EstimatorType.Uncorrelated, np.arange(64.))
    with pytest.raises(_capi.McigError, match="no CUDA device"):
        mcig.measure_peaks()


def test_state_placement_follows_the_walker_size(mcig):
    """One chain per thread keeps the walker on chip while it fits (registers, then shared memory); beyond 227 KiB per warp of
    walkers the state moves to global memory automatically. Forcing shared memory for such a walker is refused loudly."""
    from mcintegratorplusplus_b200._capi import McigError

    def make(ndim, placement=None):
        mci = mcig.MCI(ndim)
        mci.addSamplingFunction(mcig.Gauss(ndim))
        mci.addObservable(mcig.X2Sum(ndim), 0, 1)
        if placement is not None:
            mci.setStatePlacement(placement)
        mci.prebuild()
        return mci.kernelSource()

    assert "walk_kernel_reg" in make(3)
    assert "walk_kernel_smem" in make(128)  # 131 KiB per warp of walkers: fits
    assert "walk_kernel_gmem" in make(512)
    assert "walk_kernel_gmem" in make(3, placement=2)
    with pytest.raises(McigError, match="shared memory"):
        make(512, placement=1)


def test_all_move_footprint_and_unrolling_rules(mcig, monkeypatch):
    """All-moves in state memory: an element-wise sampling function with protoElement needs no proto-value arrays (views over x and
    the proposal, MS_MAIN_VPO; MCIG_ALL_VPO=0 keeps the arrays), which doubles the walkers per block; coordinate loops of 65 .. 256
    coordinates are fully unrolled (the engine prepends MCIG_UNROLL_MAX 256 unless the experiment knob names it)."""
    import re

    def make(ndim, pdf=None, placement=1):
        mci = mcig.MCI(ndim)
        mci.setRngMode(0)
        mci.setNWalkers(4096)
        mci.addSamplingFunction((pdf or mcig.ExpNDPDF)(ndim))
        mci.addObservable(mcig.XND(ndim), 20, 1)
        if placement is not None:
            mci.setStatePlacement(placement)
        mci.prebuild()
        return mci.kernelSource()

    def block(src):
        return int(re.search(r"BLOCK = (\d+)", src).group(1))

    on = make(64)
    assert "MS_MAIN_VPO = true" in on and "walk_kernel_smem" in on and "#define MCIG_UNROLL_MAX" not in on
    monkeypatch.setenv("MCIG_ALL_VPO", "0")
    off = make(64)
    assert "MS_MAIN_VPO = false" in off and block(on) == 2*block(off)
    monkeypatch.delenv("MCIG_ALL_VPO")
    assert "MS_MAIN_VPO = false" in make(8, placement=None)  # register-resident walkers keep their proto values in registers
    # automatic placement: eligible all-moves (uniform proposal, SUM_ACCEPTANCE sampling function, element-wise observables) spread one
    # walker over several lanes from 16 coordinates on; below, and for ineligible configurations, the rules above apply
    auto64 = make(64, placement=None)
    assert "walk_kernel_lanes" in auto64 and "LANES = 4, NL = 16" in auto64
    assert "LANES = 8, NL = 16" in make(128, placement=None) and "LANES = 16, NL = 16" in make(256, placement=None)
    assert "LANES = 2, NL = 16" in make(32, placement=None) and "LANES = 2, NL = 24" in make(48, placement=None) and "LANES = 2, NL = 8" in make(16, placement=None)
    assert "walk_kernel_reg" in make(8, placement=None) and "walk_kernel_smem" in make(20, placement=None)  # (20 does not split into lanes of 4 k coordinates)
    assert "LANES = 2, NL = 4" in make(8, placement=3)
    assert "#define MCIG_UNROLL_MAX 256" in make(96)
    assert "#define MCIG_UNROLL_MAX" not in make(320, placement=2)
    monkeypatch.setenv("MCIG_JIT_DEFINES", "MCIG_UNROLL_MAX=16")
    assert make(96).count("#define MCIG_UNROLL_MAX") == 1


def test_parameterised_move_distributions_host_rules(mcig):
    """mcig_set_srrd_params (the reference's `rdist` constructor argument, include/mci/SRRDAllMove.hpp:45-58): argument checks mirror the kinds, the
    Philox-mode kernel receives the parameters as JIT defines (fixed number of uniforms per value), the replay-mode kernel consumes outputs and is
    the same kernel whatever the parameters; shapes without a fixed-count sampler are refused when the kernel is generated."""
    from mcintegratorplusplus_b200._capi import McigError

    def make(mode, srrd, par, veclen=0):
        mci = mcig.MCI(4)
        mci.setRngMode(mode)
        mci.addSamplingFunction(mcig.Gauss(4))
        mci.addObservable(mcig.XND(4), 0, 1)
        mci.setTrialMove(mcig.SRRDType(srrd), veclen, params=par)
        mci.prebuild()
        return mci.kernelSource()

    src = make(0, 2, (2.0,))
    assert "#define MCIG_SRRD_PARAM 1" in src and "#define MCIG_SRRD_PAR0 0x1p+1" in src and "#define MCIG_SRRD_NU 2" in src
    assert "#define MCIG_SRRD_NU 5" in make(0, 5, (2.5, 0.5), veclen=2) and "#define MCIG_SRRD_K2A 5" in make(0, 5, (2.5, 0.5))  # 2 exponentials + 2 (half a squared normal) + sign
    assert "#define MCIG_SRRD_K2A 4" in make(0, 9, (4.0, 6.0)) and "#define MCIG_SRRD_K2B 6" in make(0, 9, (4.0, 6.0))
    assert "MCIG_SRRD_PARAM" not in make(0, 2, None) and "MCIG_SRRD_PARAM" not in make(2, 2, (2.0,))
    assert make(2, 5, (2.3, 0.7)) == make(2, 5, None)  # replay: any parameters, same kernel
    # shapes without a closed form: Marsaglia & Tsang with 6 tries of 3 uniforms (+ 1 below shape 1), + the sign
    g = make(0, 5, (2.3, 0.5))
    assert "#define MCIG_SRRD_GENA 1" in g and "#define MCIG_SRRD_GENB 0" in g and "#define MCIG_SRRD_NU 19" in g and "#define MCIG_SRRD_SHA 0x1.2666666666666p+1" in g
    assert "#define MCIG_SRRD_NU 20" in make(0, 5, (0.3, 1.0)) and "#define MCIG_SRRD_NU 19" in make(0, 8, (2.5,))  # Chisq(2.5) = 2 Gamma(1.25)
    f = make(0, 9, (3.0, 200.0))  # Fisher: Gamma(3/2) by the closed form (1 + 2 uniforms), Gamma(100) by the test
    assert "#define MCIG_SRRD_GENA 0" in f and "#define MCIG_SRRD_GENB 1" in f and "#define MCIG_SRRD_K2A 3" in f and "#define MCIG_SRRD_NU 22" in f
    assert "MCIG_SRRD_GENA" not in make(0, 5, (2.5, 0.5))
    for srrd, par, msg in ((5, (2.0,), "takes 2 parameter"), (2, (1.0, 2.0), "takes 1 parameter"), (0, (1.0,), "takes 0 parameter"), (3, (0.0,), "positive"),
                           (7, (0.1, -1.0), "positive")):
        with pytest.raises(McigError, match=msg):
            make(0, srrd, par)
    make(0, 7, (-2.0, 0.5))  # lognormal m: any real


def test_user_domain_functor_host_rules(mcig):
    """mcig_set_domain_plugin: the functor is wrapped behind the kernels' domain interface, its parameters travel in the parameter blob, the host-side
    rules take the stated sizes / volume (no sampling function: finite volume required); lane-split walkers keep to the built-in domains."""
    from mcintegratorplusplus_b200._capi import McigError
    src = """struct HalfLine { static constexpr int NPAR = 1; const double * par;
      __device__ void wrap(int, double & x) const { if (x < par[0]) { x = 2.*par[0] - x; } }
      __device__ double scale(int, double u) const { return par[0] + u/(1. - u); } };"""
    mcig.register_plugin(3, "HalfLine", "HalfLine", src, ndim=0, nvalues=0, npar=1)
    mci = mcig.MCI(64)
    mci.setRngMode(0)
    mci.addSamplingFunction(mcig.ExpNDPDF(64))
    mci.addObservable(mcig.XND(64), 20, 1)
    mci.prebuild()
    assert "walk_kernel_lanes" in mci.kernelSource()
    mci.setDomain(mcig.Domain("HalfLine", (0.5,)), 2*3.4028234663852886e+38, 0.0)
    mci.prebuild()
    s = mci.kernelSource()
    assert "typedef UserDomain<HalfLine> Domain;" in s and "struct HalfLine" in s and "walk_kernel_lanes" not in s
    with pytest.raises(McigError, match="wrong number of functor parameters"):
        mci.setDomain(mcig.Domain("HalfLine", ()), 1.0, 0.0)
    with pytest.raises(McigError, match="sizes must be positive"):
        mci.setDomain(mcig.Domain("HalfLine", (0.5,)), 0.0, 0.0)
    mci.clearSamplingFunctions()
    with pytest.raises(McigError, match="infinite domain requires a sampling function"):
        mci.integrate(100, False, False)
    mci.resetDomain()
    mci.addSamplingFunction(mcig.ExpNDPDF(64))
    mci.prebuild()
    assert "UnboundDomain Domain;" in mci.kernelSource()


def test_multistep_cold_position_rules(mcig):
    """MultiStepMove from 32 coordinates on (element-wise main and sub sampling functions): only the sub-walk's array in shared memory, block size chosen
    so that the walkers are one full round; below (and with the measurement knob MCIG_MS_COLD_X=0, read once per process) the committed position stays on chip with the 128-thread rule."""
    def make(ndim, sub=True, w=65536):
        mci = mcig.MCI(ndim)
        mci.setRngMode(0)
        mci.setNWalkers(w)
        mci.addSamplingFunction(mcig.Gauss(ndim))
        mci.addObservable(mcig.XND(ndim), 20, 1)
        mci.setTrialMove(mcig.MoveType.MultiStep, 1, nsteps=ndim, sub_pdfs=[mcig.ExpNDPDF(ndim)] if sub else [])
        mci.prebuild()
        return mci.kernelSource()

    s64 = make(64)
    assert "MS_COLD_X = true" in s64 and "MS_SUB_SUM = true" in s64 and "BLOCK = 224" in s64 and "GmemStore<64>" in s64
    s32 = make(32)
    assert "MS_COLD_X = true" in s32 and "RegStore<32>" in s32  # 32 sums stay in registers
    assert "MS_COLD_X = true" in make(64, sub=False) and "MS_COLD_X = false" in make(16)
    assert "MS_COLD_X = true" in make(64, w=4096)  # (small jobs: whatever block size fills the machine best; must compile)


def test_multistep_quad_draw_groups_host_rules(mcig, monkeypatch):
    """MultiStepMove in the Philox modes packs four sub-steps into one draw group; the HOST decides (groups_per_step() must agree with the kernels), so
    the switch travels inside the generated source: on in both Philox modes, off in replay mode (the reference's draws are consumed one by one), off
    with the measurement knob and with the pairwise sub-step experiment (which draws per sub-step). Other moves carry no switch at all."""
    def make(mode, move=None, ndim=6):
        mci = mcig.MCI(ndim)
        mci.setRngMode(mode)
        mci.addSamplingFunction(mcig.Gauss(ndim))
        mci.addObservable(mcig.XND(ndim), 0, 1)
        if move is None:
            mci.setTrialMove(mcig.MoveType.MultiStep, 1, sub_pdfs=[mcig.ExpNDPDF(ndim)])
        else:
            mci.setTrialMove(move)
        mci.prebuild()
        return mci.kernelSource()

    assert "#define MCIG_MS_QUADS 1" in make(0) and "#define MCIG_MS_QUADS 1" in make(1)
    assert "#define MCIG_MS_QUADS 0" in make(2)
    assert "MCIG_MS_QUADS" not in make(0, mcig.MoveType.All) and "MCIG_MS_QUADS" not in make(0, mcig.MoveType.Vec)
    monkeypatch.setenv("MCIG_MS_QUADS", "0")
    assert "#define MCIG_MS_QUADS 0" in make(0)
    monkeypatch.delenv("MCIG_MS_QUADS")
    monkeypatch.setenv("MCIG_JIT_DEFINES", "MCIG_MS_PAIR=1")
    assert "#define MCIG_MS_QUADS 0" in make(0, ndim=24)
