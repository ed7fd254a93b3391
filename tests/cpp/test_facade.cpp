// C++ facade test, written like the reference's own unit tests (test/ut2/main.cpp, test/ut4/main.cpp, test/ut5/main.cpp):
// plain asserts on 2-3 sigma agreement with known expectations, plus API behaviour (ownership, exceptions, getters).
// argv[1] == "api": only the checks that need no GPU.
#include <cassert>
#include <cmath>
#include <cstring>
#include <iostream>
#include <memory>
#include <stdexcept>
#include <vector>

#include "mci/DeviceFunctions.hpp"
#include "mci/Estimators.hpp"
#include "mci/MCIntegrator.hpp"
#include "mci/MPIMCI.hpp"

using namespace mci;

// A user-defined dependent observable (reference: a class deriving from ObservableFunctionInterface AND
// DependentObservableInterface): local "energy" = sum of the pdf's proto values, plus a product with observable 0.
class ProtoSumObs final: public ObservableFunctionInterface, public DependentObservableInterface
{
protected:
    ObservableFunctionInterface * _clone() const final { return new ProtoSumObs(_ndim); }

public:
    explicit ProtoSumObs(int ndim): ObservableFunctionInterface(ndim, 2, false), DependentObservableInterface(true) {}
    DeviceFunctor deviceFunctor() const final
    {
        return DeviceFunctor("ProtoSumObs", "user::ProtoSumObs<{ndim}>", R"(
namespace user {
template <int NDIM> struct ProtoSumObs {
    const double * par;
    template <class XV, class DEP> __device__ void observableFunction(const XV & x, double * out, const DEP & dep) const
    {
        double s = 0.;
        for (int i = 0; i < NDIM; ++i) { s += dep.proto(i); }
        out[0] = s;
        out[1] = x[0]*x[0] - dep.obs(0, 0)/1.; // observable 0 is X2Sum: x0^2 - sum x^2
    }
};
})");
    }
};

// A user-defined step callback: per-walker count of calls and of accepted moves
class CountingCallback final: public StepCallbackInterface
{
protected:
    StepCallbackInterface * _clone() const final { return new CountingCallback(); }

public:
    int64_t bufferDoubles(int64_t nwalkers) const final { return 2*nwalkers; }
    DeviceFunctor deviceFunctor() const final
    {
        return DeviceFunctor("CountingCallback", "user::CountingCallback", R"(
namespace user {
struct CountingCallback {
    const double * par;
    template <class XO, class XN> __device__ void operator()(const XO &, const XN &, bool accepted, long long walker, long long step, double * buf) const
    {
        if (step < 0) { return; } // the call from initializeSampling
        buf[2*walker] += 1.;
        buf[2*walker + 1] += accepted ? 1. : 0.;
    }
};
})");
    }
};

// A user-defined domain (the reference's DomainInterface is user-subclassable, include/mci/DomainInterface.hpp:26-55): a box with REFLECTING walls
// [lo, hi]^ndim. Host methods as in the reference; deviceFunctor() names the device twin of applyDomain / scaleToDomain.
class ReflectingBox final: public DomainInterface
{
    const double _lo, _hi;

protected:
    DomainInterface * _clone() const final { return new ReflectingBox(ndim, _lo, _hi); }

public:
    ReflectingBox(int n_dim, double lo, double hi): DomainInterface(n_dim), _lo(lo), _hi(hi) {}
    void applyDomain(double x[]) const final
    {
        for (int i = 0; i < ndim; ++i) {
            while (x[i] < _lo || x[i] > _hi) { x[i] = (x[i] < _lo) ? 2.*_lo - x[i] : 2.*_hi - x[i]; }
        }
    }
    void scaleToDomain(double normX[]) const final
    {
        for (int i = 0; i < ndim; ++i) { normX[i] = _lo + normX[i]*(_hi - _lo); }
    }
    void getSizes(double dimSizes[]) const final { std::fill(dimSizes, dimSizes + ndim, _hi - _lo); }
    double getVolume() const final { return pow(_hi - _lo, ndim); }
    DeviceFunctor deviceFunctor() const final
    {
        return DeviceFunctor("ReflectingBox", "user::ReflectingBox", R"(
namespace user {
struct ReflectingBox {
    static constexpr int NPAR = 2;
    const double * par; // lo, hi
    __device__ void wrap(int, double & x) const
    {
        while (x < par[0] || x > par[1]) { x = (x < par[0]) ? 2.*par[0] - x : 2.*par[1] - x; }
    }
    __device__ double scale(int, double u01) const { return par[0] + u01*(par[1] - par[0]); }
};
})", {_lo, _hi});
    }
};

// A user-defined trial move (reference: a TrialMoveInterface subclass with a host trialMove(WalkerState &, ...)): drifted uniform proposal with typed step
// sizes, acceptance factor 1 where the reverse move lies inside the proposal interval and 0 where it does not.
class DriftMove final: public TypedMoveInterface
{
    const double _drift;

protected:
    TrialMoveInterface * _clone() const final { return new DriftMove(*this); }

public:
    DriftMove(int ndim, double initStepSize, double drift): TypedMoveInterface(ndim, 1, nullptr, initStepSize, SRRDType::Uniform), _drift(drift) {}
    MoveType getMoveType() const final { return MoveType::All; }
    double getChangeRate() const final { return 1.; }
    DeviceFunctor deviceFunctor() const final
    {
        return DeviceFunctor("DriftMove", "user::DriftMove<{ndim}>", R"(
namespace user {
template <int NDIM> struct DriftMove {
    static constexpr int NPAR = 1;
    const double * par;
    template <class XO, class XN, class T, class U>
    __device__ double trialMove(const XO & xold, XN & xnew, const double * steps, T typeOf, const U & u) const
    {
        double macc = 1.;
        for (int i = 0; i < NDIM; ++i) {
            const double s = steps[typeOf.of(i)];
            const double d = s*(2.*u(i) - 1.) + par[0];
            xnew[i] = xold[i] + d;
            if (fabs(d + par[0]) > s) { macc = 0.; }
        }
        return macc;
    }
};
})", {_drift});
    }
};

template <class E, class F>
static bool throws(F f)
{
    try { f(); }
    catch (const E &) { return true; }
    catch (...) { return false; }
    return false;
}

static void test_api()
{
    MCI mci(3);
    assert(mci.getNDim() == 3 && mci.getNObs() == 0 && mci.getNPDF() == 0);
    assert(mci.getTargetAcceptanceRate() == 0.5 && mci.getNfindMRT2Iterations() == -50 && mci.getNdecorrelationSteps() == -10000);
    assert(mci.getMRT2Step(0) == 0.05 && mci.getMRT2Step(1) == 0.); // DEFAULT_MRT2STEP, out-of-range index -> 0
    assert(!mci.getDomain().isFinite());
    // setters / getters (test/ut2/main.cpp:20-54)
    double x[3] = {5., -5., 10.};
    mci.setX(x);
    assert(mci.getX(0) == 5. && mci.getX(1) == -5. && mci.getX(2) == 10.);
    assert(mci.getX()[2] == 10.);
    mci.setX(1, 2.5);
    assert(mci.getX(1) == 2.5);
    mci.setIRange(-1., 1.); // periodic wrap of the current position
    assert(mci.getDomain().isFinite() && mci.getDomain().getVolume() == 8.);
    for (int i = 0; i < 3; ++i) { assert(mci.getX(i) >= -1. && mci.getX(i) <= 1.); }
    mci.newRandomX();
    for (int i = 0; i < 3; ++i) { assert(mci.getX(i) >= -1. && mci.getX(i) <= 1.); }
    auto olddom = mci.resetDomain();
    assert(olddom->isFinite() && !mci.getDomain().isFinite());
    // ownership: set*/pop* hand the previous object back
    auto oldmove = mci.setTrialMove(MoveType::Vec);
    assert(oldmove->getMoveType() == MoveType::All && mci.getTrialMove().getChangeRate() == 1./3.);
    int typeEnds[2] = {1, 3};
    mci.setTrialMove(SRRDType::Uniform, 0, 2, typeEnds);
    assert(mci.getTrialMove().getNStepSizes() == 2 && mci.getTrialMove().getStepSizeIndex(0) == 0 && mci.getTrialMove().getStepSizeIndex(2) == 1);
    double steps[2] = {0.3, 0.6};
    mci.setMRT2Step(steps);
    assert(mci.getMRT2Step(1) == 0.6);
    mci.addSamplingFunction(ThreeDimGaussianPDF());
    mci.addObservable(XSquared());
    mci.addObservable(XYZSquared(), 16, 1);
    assert(mci.getNObs() == 2 && mci.getNObsDim() == 4 && mci.getNPDF() == 1);
    auto popped = mci.popObservable();
    assert(popped->getNObs() == 3 && mci.getNObsDim() == 1);
    // exceptions
    assert(throws<std::invalid_argument>([&] { mci.addSamplingFunction(Exp1DPDF()); }));
    assert(throws<std::invalid_argument>([&] { mci.addObservable(X1D()); }));
    assert(throws<std::invalid_argument>([&] { mci.addObservable(XSquared(), 1, 1, true, EstimatorType::Noop); }));
    assert(throws<std::invalid_argument>([&] { mci.setTrialMove(SRRDType::Uniform, 2); }));
    assert(throws<std::invalid_argument>([&] { mci.setIRange(1., -1.); }));
    assert(throws<std::invalid_argument>([&] { mci.setDomain(OrthoPeriodicDomain(2, -1., 1.)); }));
    mci.setTrialMove(SRRDType::Cauchy, 1);
    assert(mci.getTrialMove().getMoveType() == MoveType::Vec && mci.getTrialMove().getSRRDType() == SRRDType::Cauchy);
    mci.setTrialMove(SRRDType::Gaussian);
    assert(mci.getTrialMove().getSRRDType() == SRRDType::Gaussian);
    assert(throws<std::invalid_argument>([] { selectEstimatorType(true, false); }));
    mci.setTrialMove(GaussianVecMove(3, 1, 0.4)); // named instantiations (include/mci/SRRDVecMove.hpp:101-111)
    assert(mci.getTrialMove().getSRRDType() == SRRDType::Gaussian && mci.getTrialMove().getMoveType() == MoveType::Vec && mci.getMRT2Step(0) == 0.4);
    mci.setTrialMove(FisherAllMove(3, 0.2));
    assert(mci.getTrialMove().getSRRDType() == SRRDType::Fisher && mci.getTrialMove().getMoveType() == MoveType::All);
    MultiStepMove msm(3);
    assert(msm.getNSteps() == 3 && msm.getChangeRate() == 1.);
    assert(throws<std::invalid_argument>([&] { msm.addSamplingFunction(Exp1DPDF()); }));
    std::cout << "api ok" << std::endl;
}

static void test_gpu()
{
    double avg[4], err[4];
    { // test/ut2/main.cpp:56-91 — bad start needs the warm-up; with calibration + decorrelation the result is right
        MCI mci(3);
        mci.setSeed(5649871);
        mci.setNWalkers(256);
        double x[3] = {5., -5., 10.};
        mci.setX(x);
        mci.addSamplingFunction(ThreeDimGaussianPDF());
        mci.addObservable(XSquared());
        mci.integrate(10000, avg, err, false, false);
        assert(fabs(avg[0] - 0.5) > 2.*err[0]); // no warm-up from (5,-5,10): biased
        mci.setX(x);
        mci.setMRT2Step(0.05);
        mci.integrate(10000, avg, err, true, true);
        assert(fabs(avg[0] - 0.5) < 3.*err[0]);
        assert(fabs(mci.getAcceptanceRate() - 0.5) < 0.05 && mci.getMRT2Step(0) > 0.5);
        // no sampling function over a box (test/ut2/main.cpp:94-108)
        mci.clearSamplingFunctions();
        mci.clearObservables();
        mci.addObservable(GaussXSquared());
        mci.setIRange(-5., 5.);
        mci.integrate(10000, avg, err);
        assert(fabs(avg[0] - 0.5) < 3.*err[0]);
        assert(throws<std::domain_error>([&] { mci.resetDomain(); mci.integrate(100, avg, err); }));
    }
    { // test/ut4/main.cpp:33-73 — fixed calibration/decorrelation, MJBlocker on 16384 samples, fixed blocks
        MCI mci(3);
        mci.setSeed(1337);
        mci.setNWalkers(128);
        mci.setNfindMRT2Iterations(20);
        mci.setNdecorrelationSteps(2000);
        mci.addSamplingFunction(ThreeDimGaussianPDF());
        mci.addObservable(XSquared(), 1, 1);
        mci.addObservable(XYZSquared(), 16, 1);
        mci.integrate(16384, avg, err);
        for (int i = 0; i < 4; ++i) { assert(err[i] > 0. && fabs(avg[i] - 0.5) < 3.5*err[i]); }
        assert(throws<std::invalid_argument>([&] { mci.integrate(1000, avg, err, false, false); })); // 1000 % 16 != 0
    }
    { // test/ut5/main.cpp:113-137 — custom MultiStepMove with sub-move UniformVecMove and sub-pdf ExpNDPDF, target 0.85
        MCI mci(3);
        mci.setSeed(1337);
        mci.setNWalkers(128);
        MultiStepMove msm(3, 3);
        msm.setTrialMove(UniformVecMove(3, 1, 0.1));
        msm.addSamplingFunction(ExpNDPDF(3));
        mci.setTrialMove(msm);
        mci.setTargetAcceptanceRate(0.85);
        mci.addSamplingFunction(Gauss(3));
        mci.addObservable(XSquared());
        mci.addObservable(X2(3), 1, 3);
        mci.integrate(32768, avg, err);
        for (int i = 0; i < 4; ++i) { assert(fabs(avg[i] - 0.5) < 3.5*err[i]); }
        assert(fabs(mci.getAcceptanceRate() - 0.85) < 0.06);
    }
    { // test/ut5/main.cpp:110-137, i == 0 — customised Student-t all-move built around a pre-made distribution object, and the uniform vec-move with one
        MCI mci(3);
        mci.setSeed(1337);
        mci.setNWalkers(128);
        auto customStudentDist = std::student_t_distribution<double>(2);
        auto defaultUniformDist = std::uniform_real_distribution<double>(-1., 1.); // because we can
        StudentAllMove customAllMove(mci.getNDim(), 0.05, &customStudentDist);
        UniformVecMove customVecMove(mci.getNDim(), 1, 0.1, &defaultUniformDist);
        assert(customAllMove.getSRRDParams().size() == 1 && customAllMove.getSRRDParams()[0] == 2. && customVecMove.getSRRDParams().empty());
        mci.setTrialMove(customAllMove);
        mci.addSamplingFunction(Gauss(3));
        mci.addObservable(XSquared(), 1, 1);
        mci.addObservable(X2(3), 1, 1);
        mci.integrate(32768, avg, err, true, false);
        for (int i = 0; i < 4; ++i) { assert(fabs(avg[i] - 0.5) < 3.5*err[i]); }
        assert(fabs(mci.getAcceptanceRate() - 0.5) < 0.05);
        // symmetrised gamma with half-integer shape (closed form), then with a general shape (fixed-count Marsaglia-Tsang); asymmetric objects are refused
        auto gam = SymmetrizedPRRD<std::gamma_distribution<double>>(std::gamma_distribution<double>(2.5, 0.5));
        mci.setTrialMove(GammaAllMove(3, 0.3, &gam));
        assert(mci.getTrialMove().getSRRDParams().size() == 2);
        mci.integrate(32768, avg, err, false, false);
        for (int i = 0; i < 4; ++i) { assert(fabs(avg[i] - 0.5) < 3.5*err[i]); }
        auto gam2 = SymmetrizedPRRD<std::gamma_distribution<double>>(std::gamma_distribution<double>(2.3, 0.5));
        mci.setTrialMove(GammaAllMove(3, 0.3, &gam2));
        mci.integrate(32768, avg, err, false, false);
        for (int i = 0; i < 4; ++i) { assert(fabs(avg[i] - 0.5) < 3.5*err[i]); }
        // the reference's vec-move clone drops the distribution (include/mci/SRRDVecMove.hpp:30-33): MCI holds a default-parameter move
        GammaVecMove gv(3, 1, 0.8, &gam);
        assert(gv.getSRRDParams().size() == 2);
        mci.setTrialMove(gv);
        assert(mci.getTrialMove().getSRRDParams().empty() && mci.getTrialMove().getSRRDType() == SRRDType::Gamma);
        auto shifted = std::normal_distribution<double>(1., 2.);
        assert(throws<std::invalid_argument>([&] { GaussianAllMove bad(3, 0.1, &shifted); }));
        auto wide = std::uniform_real_distribution<double>(-2., 2.);
        assert(throws<std::invalid_argument>([&] { UniformAllMove bad(3, 0.1, &wide); }));
    }
    { // user-defined trial move: detailed balance through its own acceptance factor, step calibrated like any typed move
        MCI mci(3);
        mci.setSeed(777);
        mci.setNWalkers(256);
        mci.setTrialMove(DriftMove(3, 0.3, 0.03));
        assert(mci.getTrialMove().getNStepSizes() == 1 && mci.getMRT2Step(0) == 0.3);
        mci.setTargetAcceptanceRate(0.4);
        mci.addSamplingFunction(Gauss(3));
        mci.addObservable(XSquared(), 1, 1);
        mci.addObservable(XND(3), 1, 1);
        mci.integrate(16384, avg, err, true, true);
        assert(fabs(avg[0] - 0.5) < 3.5*err[0]);
        for (int i = 1; i < 4; ++i) { assert(fabs(avg[i]) < 3.5*err[i]); } // <x_i> = 0: a drifted proposal WITHOUT its acceptance factor would bias this
        assert(fabs(mci.getAcceptanceRate() - 0.4) < 0.06 && mci.getMRT2Step(0) > 0.3);
    }
    { // user-defined domain: exp(-r^2) restricted to the reflecting box [-1, 1.5]^3, and plain sampling of the box without a sampling function
        MCI mci(3);
        mci.setSeed(4242);
        mci.setNWalkers(512);
        mci.setDomain(ReflectingBox(3, -1., 1.5));
        const double far[3] = {3.2, -2.6, 0.4};
        mci.setX(far); // reflected into the box by the host twin: 3.2 -> -0.2, -2.6 -> 0.6
        assert(fabs(mci.getX(0) + 0.2) < 1e-12 && fabs(mci.getX(1) - 0.6) < 1e-12 && mci.getX(2) == 0.4);
        mci.addSamplingFunction(Gauss(3));
        mci.addObservable(XND(3), 1, 1);
        mci.integrate(8192, avg, err, true, true);
        // <x> of exp(-x^2) on [-1, 1.5]: (e^{-1} - e^{-2.25})/2 / (sqrt(pi)/2 (erf(1) + erf(1.5)))
        const double want = 0.5*(exp(-1.) - exp(-2.25))/(0.5*sqrt(M_PI)*(erf(1.) + erf(1.5)));
        for (int i = 0; i < 3; ++i) { assert(err[i] > 0. && fabs(avg[i] - want) < 4.*err[i]); }
        assert(mci.getMRT2Step(0) <= 0.5*2.5); // calibrated step capped at half the domain size (src/MCIntegrator.cpp:151-155)
        for (int i = 0; i < 3; ++i) { assert(mci.getX(i) >= -1. && mci.getX(i) <= 1.5); }
        mci.clearSamplingFunctions();
        mci.clearObservables();
        mci.addObservable(XND(3), 1, 1);
        mci.integrate(8192, avg, err, false, false); // no pdf: uniform over the box, result times the volume
        for (int i = 0; i < 3; ++i) { assert(fabs(avg[i] - 0.25*pow(2.5, 3)) < 4.*err[i]); }
    }
    { // replay mode, 1 walker: bit-exact reference numbers (generated from the compiled reference)
        MCI mci(3);
        mci.setRngMode(RngMode::Replay);
        mci.setSeed(5649871);
        mci.addSamplingFunction(ThreeDimGaussianPDF());
        mci.addObservable(XSquared(), 0, 1);
        mci.setMRT2Step(1.0);
        mci.integrate(100000, avg, err, false, false);
        assert(fabs(avg[0] - 0.49129926481208264) < 1e-12*0.5);
        assert(mci.getAcceptanceRate() == 0.50334);
        assert(mci.getX(0) == 0.56622698245761904 && mci.getX(2) == 0.65777044965711584);
    }
    { // dependent observable + step callback as device functors
        MCI mci(3);
        mci.setSeed(99);
        mci.setNWalkers(512);
        mci.setMRT2Step(0.9);
        mci.addSamplingFunction(Gauss(3));
        mci.addObservable(X2Sum(3), 1, 1);
        mci.addObservable(ProtoSumObs(3), 1, 1);
        mci.setCallback(CountingCallback());
        assert(throws<std::logic_error>([&] { mci.setCallback([](const MCI &) {}); }));
        double a[3], e[3];
        mci.integrate(4096, a, e, false, false);
        assert(fabs(a[0] - 1.5) < 4.*e[0]);             // <sum x^2> = 3 * 0.5
        assert(fabs(a[1] - a[0]) < 1e-12);               // Gauss proto values are x_i^2: their sum is X2Sum, sample by sample
        assert(fabs(a[2] + 1.0) < 4.*e[2]);              // <x0^2 - sum x^2> = -1
        const std::vector<double> buf = mci.getCallbackBuffer();
        assert(buf.size() == 1024);
        double calls = 0., acc = 0.;
        for (size_t w = 0; w < 512; ++w) { calls += buf[2*w]; acc += buf[2*w + 1]; }
        assert(calls == 512.*4096. && fabs(acc/calls - mci.getAcceptanceRate()) < 1e-12);
        mci.clearCallback();
        mci.integrate(4096, a, e, false, false);
    }
    { // free estimator functions on host data
        std::vector<double> x(4096);
        for (size_t i = 0; i < x.size(); ++i) { x[i] = sin(0.37*static_cast<double>(i)) + 0.001*static_cast<double>(i % 7); }
        double a1, e1, a2[1], e2[1];
        OneDimUncorrelatedEstimator(4096, x.data(), a1, e1);
        MJBlockerEstimator(4096, 1, x.data(), a2, e2);
        assert(fabs(a1 - a2[0]) < 1e-14 && e1 > 0. && e2[0] > 0.);
        assert(throws<std::invalid_argument>([&] { MJBlockerEstimator(1000, 1, x.data(), a2, e2); }));
        double a3, e3;
        OneDimBlockEstimator(4096, x.data(), 16, a3, e3); // 16 blocks of 256 samples
        assert(fabs(a3 - a1) < 1e-14 && e3 > 0.);
        assert(throws<std::invalid_argument>([&] { OneDimBlockEstimator(10, x.data(), 11, a3, e3); }));
    }
    std::cout << "gpu ok" << std::endl;
}

int main(int argc, char ** argv)
{
    test_api();
    if (argc > 1 && strcmp(argv[1], "api") == 0) { return 0; }
    test_gpu();
    return 0;
}
