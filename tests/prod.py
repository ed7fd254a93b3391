"""Build a product MCI (CUDA path through the C-ABI) from the same keyword spec the oracle configs use (tests/configs.py)."""
import orc

PDF_NAMES = {orc.PDF_GAUSS3D: "ThreeDimGaussianPDF", orc.PDF_GAUSS: "Gauss", orc.PDF_EXP1D: "Exp1DPDF", orc.PDF_EXPND: "ExpNDPDF",
             orc.PDF_NORMLINE: "NormalizedLine"}
OBS_NAMES = {orc.OBS_XSQUARED: "XSquared", orc.OBS_GAUSSXSQUARED: "GaussXSquared", orc.OBS_XYZSQUARED: "XYZSquared", orc.OBS_X1D: "X1D",
             orc.OBS_XND: "XND", orc.OBS_UPDXND: "UpdateableXND", orc.OBS_CONSTVAL: "Constval", orc.OBS_POLYNOM: "Polynom",
             orc.OBS_X2SUM: "X2Sum", orc.OBS_X2: "X2", orc.OBS_PARABOLA: "Parabola", orc.OBS_NORMPARABOLA: "NormalizedParabola",
             orc.OBS_DEPENDENT: "HarnessDepObs"}

# Device twins of the reference-side test classes in oracle/ref_harness.cpp (user plugins: CUDA C++ source compiled by NVRTC)
DEP_OBS_SRC = """
namespace test_plugins {
template <int NDIM>
struct HarnessDepObs { // DependentObservableInterface: sum of the pdf's proto values + observable 0, and x[0]*observable 0
    const double * par;
    template <class XV, class DEP>
    __device__ void observableFunction(const XV & in, double * out, const DEP & dep) const
    {
        double s = 0.;
        for (int i = 0; i < NDIM; ++i) { s += dep.proto(i); } // Gauss(ndim): one proto value per coordinate
        out[0] = s + dep.obs(0, 0);
        out[1] = in[0]*dep.obs(0, 0);
    }
};
}
"""
CALLBACK_SRC = """
namespace test_plugins {
template <int NDIM>
struct FoldCallback { // the four sums of mciref_run_callback, one set per walker
    const double * par;
    template <class XO, class XN>
    __device__ void operator()(const XO & xold, const XN & xnew, bool accepted, long long walker, long long step, double * buf) const
    {
        double * b = buf + 6*walker;
        b[0] += 1.;
        b[1] += accepted ? 1. : 0.;
        b[2] += accepted ? xnew[0] : xold[0];
        b[3] += xnew[NDIM - 1] - xold[NDIM - 1];
        b[4] += (double)step;
        b[5] = (double)step;
    }
};
}
"""


DRIFT_MOVE_SRC = """namespace test_plugins {
// device twin of oracle/ref_harness.cpp: HarnessDriftMove (a user-defined TrialMoveInterface subclass): x' = x + step (2u - 1) + drift per coordinate,
// acceptance factor 1 where the reverse move lies inside the proposal interval, else 0
template <int NDIM> struct DriftMove { static constexpr int NPAR = 1; const double * par;
  template <class XO, class XN, class T, class U>
  __device__ double trialMove(const XO & xold, XN & xnew, const double * steps, T typeOf, const U & u) const {
    double macc = 1.;
    for (int i = 0; i < NDIM; ++i) {
      const double s = steps[typeOf.of(i)];
      const double d = s*(2.*u(i) - 1.) + par[0];
      xnew[i] = xold[i] + d;
      if (fabs(d + par[0]) > s) { macc = 0.; }
    }
    return macc;
  } };
}"""


def register_test_plugins(m):
    m.register_plugin(4, "DriftMove", "test_plugins::DriftMove<{ndim}>", DRIFT_MOVE_SRC, ndim=0, nvalues=0, npar=1)
    m.register_plugin(1, "HarnessDepObs", "test_plugins::HarnessDepObs<{ndim}>", DEP_OBS_SRC, ndim=0, nvalues=2, npar=0, dependent=True)
    m.register_plugin(2, "FoldCallback", "test_plugins::FoldCallback<{ndim}>", CALLBACK_SRC, ndim=0, npar=0)



def build_mci(m, spec, nwalkers=1, mode=None, seeds=None, placement=None):
    kw = dict(spec)
    ndim = kw["ndim"]
    register_test_plugins(m)
    mci = m.MCI(ndim)
    mci.setRngMode(m.RngMode.Replay if mode is None else mode)
    if nwalkers != 1:
        mci.setNWalkers(nwalkers)
    mci.setSeed(kw["seed"])
    if seeds is not None:
        mci.setWalkerSeeds(seeds)
    if kw.get("lb") is not None:
        mci.setIRange(kw["lb"], kw["ub"])
    mt = kw.get("move_type", orc.MOVE_ALL)
    ntypes = kw.get("ntypes", 1)
    te = kw.get("type_ends")
    srrd = m.SRRDType(kw.get("srrd", 0))
    par = kw.get("srrd_par") or None
    if mt == orc.MOVE_USER_DRIFT:
        mci.setTrialMove(m.Move("DriftMove", kw["srrd_par"]), 0, ntypes, te)
    elif mt == orc.MOVE_ALL:
        mci.setTrialMove(srrd, 0, ntypes, te, params=par)
    elif mt == orc.MOVE_VEC:
        # the reference drops a vec-move's pre-made distribution when MCI clones the move (include/mci/SRRDVecMove.hpp:30-33): the goldens of the
        # par_*_vec runs are those of the DEFAULT distribution, and the C++ facade mirrors that (include/mci/TrialMoveInterface.hpp: SRRDVecMove::_clone)
        mci.setTrialMove(srrd, max(1, kw.get("veclen", 1)), ntypes, te, params=None)
    else:
        sub = []
        if kw.get("ms_sub_pdf_id", 0):
            sub.append(m.SamplingFunction(PDF_NAMES[kw["ms_sub_pdf_id"]]))
        mci.setTrialMove(m.MoveType.MultiStep, max(1, kw.get("veclen", 1)), ntypes, te, nsteps=kw.get("ms_nsteps", 0), sub_pdfs=sub)
    steps = list(kw.get("steps", (0.05,)))
    for i in range(max(1, ntypes)):
        mci.setMRT2Step(i, steps[i] if i < len(steps) else steps[-1])
    mci.setX(kw.get("x0") if kw.get("x0") is not None else [0.0]*ndim)
    if kw["pdf_id"]:
        mci.addSamplingFunction(m.SamplingFunction(PDF_NAMES[kw["pdf_id"]]))
    for o in kw["obs"]:
        o = tuple(o)
        bs, ns = max(0, o[1]), max(1, o[2])
        fe = o[3] if len(o) > 3 else bs > 0
        et = o[4] if len(o) > 4 else orc.default_estim(bs)
        mci.addObservable(m.Observable(OBS_NAMES[o[0]]), bs, ns, fe, m.EstimatorType(et))
    mci.setTargetAcceptanceRate(kw.get("target_acc", 0.5))
    mci.setNfindMRT2Iterations(kw.get("nfind", -50))
    mci.setNdecorrelationSteps(kw.get("ndecorr", -10000))
    if placement is not None:
        mci.setStatePlacement(placement)
    return mci
