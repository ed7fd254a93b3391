// ref_harness.cpp — TEST INFRASTRUCTURE ONLY.
//
// Thin C-ABI driver around the UNMODIFIED reference library (DCM-UPB/MCIntegratorPlusPlus). It is compiled
// by oracle/Makefile together with the reference's own sources *where they lie* under /root/reference
// (nothing is copied into this repo) into oracle/_ref/libmci_ref.so. It exists to
//   (1) pin the plain-C restatement (oracle/mci_oracle.c) against the real implementation, bit for bit, and
//   (2) generate the golden vectors committed under tests/golden/ (tools: oracle/gen_golden.py).
// The product library never links, loads or calls this file.
//
// All std headers are included BEFORE the `private -> public` switch, which is applied only to the
// reference's own headers so that the harness can read MCI's walker state / RNG inside the step callback
// (access specifiers do not change class layout, the library itself is compiled untouched).
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <fstream>
#include <functional>
#include <iostream>
#include <limits>
#include <memory>
#include <numeric>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>
#include <array>

#define private public
#define protected public
#include "mci/MCIntegrator.hpp"
#include "mci/Estimators.hpp"
#include "mci/MJBlocker.hpp"
#include "mci/OrthoPeriodicDomain.hpp"
#include "mci/UnboundDomain.hpp"
#include "TestMCIFunctions.hpp"   // /root/reference/test/common
#include "ExampleFunctions.hpp"   // /root/reference/examples/common
#undef private
#undef protected

#include "oracle_config.h"

using namespace mci;

static thread_local std::string g_err;

static std::unique_ptr<SamplingFunctionInterface> make_pdf(int id, int ndim)
{
    switch (id) {
    case ORC_PDF_GAUSS3D: return std::make_unique<ThreeDimGaussianPDF>();
    case ORC_PDF_GAUSS: return std::make_unique<Gauss>(ndim);
    case ORC_PDF_EXP1D: return std::make_unique<Exp1DPDF>();
    case ORC_PDF_EXPND: return std::make_unique<ExpNDPDF>(ndim);
    case ORC_PDF_NORMLINE: return std::make_unique<NormalizedLine>();
    default: throw std::invalid_argument("ref_harness: unknown pdf id");
    }
}

// A dependent observable (include/mci/DependentObservableInterface.hpp) for the golden vectors of the device-side replacement:
// depends on the sampling function (sum of the first pdf's proto values at the current position) and on observable 0.
//   out[0] = sum_i protoold_i + obs0[0]        out[1] = x[0]*obs0[0]
class HarnessDepObs final: public ObservableFunctionInterface, public DependentObservableInterface
{
    const SamplingFunctionInterface * _pdf = nullptr;
    const AccumulatorInterface * _dep = nullptr;

protected:
    ObservableFunctionInterface * _clone() const final { return new HarnessDepObs(_ndim); }

public:
    explicit HarnessDepObs(int ndim): ObservableFunctionInterface(ndim, 2, false), DependentObservableInterface(true) {}
    void registerDeps(const SamplingFunctionContainer &pdfcont, const std::vector<AccumulatorInterface *> &accuvec, int selfIdx) final
    {
        if (selfIdx < 1 || !isObsDepValid(accuvec, selfIdx, 0)) { throw std::runtime_error("HarnessDepObs: invalid dependency"); }
        _pdf = &pdfcont.getSamplingFunction(0);
        _dep = accuvec[0];
    }
    void deregisterDeps() final { _pdf = nullptr; _dep = nullptr; }
    void observableFunction(const double in[], double out[]) final
    {
        double s = 0.;
        for (int i = 0; i < _pdf->getNProto(); ++i) { s += _pdf->_protoold[i]; }
        out[0] = s + _dep->getObsValue(0);
        out[1] = in[0]*_dep->getObsValue(0);
    }
};

static std::unique_ptr<ObservableFunctionInterface> make_obs(int id, int ndim)
{
    switch (id) {
    case ORC_OBS_XSQUARED: return std::make_unique<XSquared>();
    case ORC_OBS_GAUSSXSQUARED: return std::make_unique<GaussXSquared>();
    case ORC_OBS_XYZSQUARED: return std::make_unique<XYZSquared>();
    case ORC_OBS_X1D: return std::make_unique<X1D>();
    case ORC_OBS_XND: return std::make_unique<XND>(ndim);
    case ORC_OBS_UPDXND: return std::make_unique<UpdateableXND>(ndim);
    case ORC_OBS_CONSTVAL: return std::make_unique<Constval>(ndim);
    case ORC_OBS_POLYNOM: return std::make_unique<Polynom>(ndim);
    case ORC_OBS_X2SUM: return std::make_unique<X2Sum>(ndim);
    case ORC_OBS_X2: return std::make_unique<X2>(ndim);
    case ORC_OBS_PARABOLA: return std::make_unique<Parabola>();
    case ORC_OBS_NORMPARABOLA: return std::make_unique<NormalizedParabola>();
    case ORC_OBS_DEPENDENT: return std::make_unique<HarnessDepObs>(ndim);
    default: throw std::invalid_argument("ref_harness: unknown obs id");
    }
}

static EstimatorType to_estim(int t)
{
    switch (t) {
    case ORC_EST_NOOP: return EstimatorType::Noop;
    case ORC_EST_UNCORRELATED: return EstimatorType::Uncorrelated;
    case ORC_EST_CORRELATED: return EstimatorType::Correlated;
    case ORC_EST_FCBLOCKER: return EstimatorType::FCBlocker;
    case ORC_EST_MJBLOCKER: return EstimatorType::MJBlocker;
    default: throw std::invalid_argument("ref_harness: unknown estimator type");
    }
}

// A user-defined trial move (include/mci/TrialMoveInterface.hpp:16-70 is meant to be subclassed) for the golden vectors of the device-side move functor:
// every coordinate moves by step*(2u - 1) + drift with u ~ U[0,1) drawn from MCI's generator. The proposal density is uniform on an interval that is NOT
// centred on the old position, so the move's acceptance factor q(x|x')/q(x'|x) is 1 where the reverse move is possible (|step v + 2 drift| <= step, per
// coordinate) and 0 where it is not: exercises pdfAcc*moveAcc (src/MCIntegrator.cpp:343) without any transcendental in the positions.
class HarnessDriftMove final: public TypedMoveInterface
{
    const double _drift;
    std::uniform_real_distribution<double> _rd;

protected:
    TrialMoveInterface * _clone() const final
    {
        auto * m = new HarnessDriftMove(_ndim, _ntypes, _typeEnds, 0., _drift);
        std::copy(_stepSizes, _stepSizes + _ntypes, m->_stepSizes);
        return m;
    }

public:
    HarnessDriftMove(int ndim, int ntypes, const int typeEnds[], double initStepSize, double drift):
            TypedMoveInterface(ndim, 0, ntypes, typeEnds, initStepSize), _drift(drift), _rd(0., 1.) {}
    double getChangeRate() const final { return 1.; }
    void protoFunction(const double[], double[]) final {}
    double trialMove(WalkerState &wlk, const double[], double[]) final
    {
        double macc = 1.;
        int xidx = 0;
        for (int tidx = 0; tidx < _ntypes; ++tidx) {
            while (xidx < _typeEnds[tidx]) {
                const double d = _stepSizes[tidx]*(2.*_rd(*_rgen) - 1.) + _drift;
                wlk.xnew[xidx] += d;
                if (fabs(d + _drift) > _stepSizes[tidx]) { macc = 0.; } // the reverse move -d - ... lies outside the proposal interval
                ++xidx;
            }
        }
        wlk.nchanged = _ndim;
        return macc;
    }
};

// A move built around a pre-made distribution, as user code does (include/mci/SRRDAllMove.hpp:45-58, SRRDVecMove.hpp:41-68, test/ut5/main.cpp:110-113)
template <class AllMoveT, class VecMoveT, class Dist>
static void set_custom_move(MCI &mci, const orc_config_t &c, int ntypes, const int * tends, const Dist &dist)
{
    if (c.move_type == ORC_MOVE_ALL) {
        AllMoveT mv(c.ndim, ntypes, ntypes > 1 ? tends : nullptr, DEFAULT_MRT2STEP, &dist);
        mci.setTrialMove(mv);
    }
    else {
        const int veclen = std::max(1, c.veclen);
        VecMoveT mv(c.ndim/veclen, veclen, ntypes, ntypes > 1 ? tends : nullptr, DEFAULT_MRT2STEP, &dist);
        mci.setTrialMove(mv);
    }
}

static void set_parameterised_move(MCI &mci, const orc_config_t &c, int ntypes, const int * tends)
{
    const double p0 = c.srrd_par[0], p1 = (c.srrd_npar > 1) ? c.srrd_par[1] : 1.;
    if (c.move_type != ORC_MOVE_ALL && c.move_type != ORC_MOVE_VEC) { throw std::invalid_argument("ref_harness: distribution parameters need an all- or vec-move"); }
    switch (c.srrd) {
    case 1: set_custom_move<GaussianAllMove, GaussianVecMove>(mci, c, ntypes, tends, std::normal_distribution<double>(0., p0)); break;
    case 2: set_custom_move<StudentAllMove, StudentVecMove>(mci, c, ntypes, tends, std::student_t_distribution<double>(p0)); break;
    case 3: set_custom_move<CauchyAllMove, CauchyVecMove>(mci, c, ntypes, tends, std::cauchy_distribution<double>(0., p0)); break;
    case 4: set_custom_move<ExponentialAllMove, ExponentialVecMove>(mci, c, ntypes, tends, SymmetrizedPRRD<std::exponential_distribution<double>>(std::exponential_distribution<double>(p0))); break;
    case 5: set_custom_move<GammaAllMove, GammaVecMove>(mci, c, ntypes, tends, SymmetrizedPRRD<std::gamma_distribution<double>>(std::gamma_distribution<double>(p0, p1))); break;
    case 6: set_custom_move<WeibullAllMove, WeibullVecMove>(mci, c, ntypes, tends, SymmetrizedPRRD<std::weibull_distribution<double>>(std::weibull_distribution<double>(p0, p1))); break;
    case 7: set_custom_move<LognormalAllMove, LognormalVecMove>(mci, c, ntypes, tends, SymmetrizedPRRD<std::lognormal_distribution<double>>(std::lognormal_distribution<double>(p0, p1))); break;
    case 8: set_custom_move<ChisqAllMove, ChisqVecMove>(mci, c, ntypes, tends, SymmetrizedPRRD<std::chi_squared_distribution<double>>(std::chi_squared_distribution<double>(p0))); break;
    case 9: set_custom_move<FisherAllMove, FisherVecMove>(mci, c, ntypes, tends, SymmetrizedPRRD<std::fisher_f_distribution<double>>(std::fisher_f_distribution<double>(p0, p1))); break;
    default: throw std::invalid_argument("ref_harness: this distribution takes no parameters");
    }
}

static void configure(MCI &mci, const orc_config_t &c)
{
    mci.setSeed(c.seed);
    if (c.domain == ORC_DOMAIN_ORTHO) { mci.setIRange(c.lb, c.ub); }
    // move
    const int ntypes = std::max(1, c.ntypes);
    std::vector<int> tends(c.type_ends, c.type_ends + ORC_MAXTYPES);
    if (c.srrd < 0 || c.srrd > 9) { throw std::invalid_argument("ref_harness: unsupported srrd"); }
    static const SRRDType kinds[10] = {SRRDType::Uniform, SRRDType::Gaussian, SRRDType::Student, SRRDType::Cauchy, SRRDType::Exponential,
                                       SRRDType::Gamma, SRRDType::Weibull, SRRDType::Lognormal, SRRDType::Chisq, SRRDType::Fisher};
    const SRRDType srrd = kinds[c.srrd];
    if (c.move_type == ORC_MOVE_USER_DRIFT) {
        HarnessDriftMove mv(c.ndim, ntypes, ntypes > 1 ? tends.data() : nullptr, DEFAULT_MRT2STEP, c.srrd_par[0]);
        mci.setTrialMove(mv);
    }
    else if (c.srrd_npar > 0) { set_parameterised_move(mci, c, ntypes, tends.data()); }
    else if (c.move_type == ORC_MOVE_ALL) {
        mci.setTrialMove(srrd, 0, ntypes, ntypes > 1 ? tends.data() : nullptr);
    }
    else if (c.move_type == ORC_MOVE_VEC) {
        mci.setTrialMove(srrd, std::max(1, c.veclen), ntypes, ntypes > 1 ? tends.data() : nullptr);
    }
    else if (c.move_type == ORC_MOVE_MULTISTEP) {
        // MultiStepMove with a uniform single-vector sub-move (its default kind, MultiStepMove.hpp:46-52)
        MultiStepMove msm(c.ndim, c.ms_nsteps > 0 ? c.ms_nsteps : c.ndim);
        const int veclen = std::max(1, c.veclen);
        if (ntypes > 1) {
            UniformVecMove sub(c.ndim/veclen, veclen, ntypes, tends.data(), DEFAULT_MRT2STEP);
            msm.setTrialMove(sub);
        }
        else {
            UniformVecMove sub(c.ndim/veclen, veclen, DEFAULT_MRT2STEP);
            msm.setTrialMove(sub);
        }
        if (c.ms_sub_pdf_id != ORC_PDF_NONE) { msm.addSamplingFunction(*make_pdf(c.ms_sub_pdf_id, c.ndim)); }
        mci.setTrialMove(msm);
    }
    else { throw std::invalid_argument("ref_harness: unknown move type"); }
    for (int i = 0; i < ntypes; ++i) { mci.setMRT2Step(i, c.steps[i]); }

    mci.setX(c.x0);
    if (c.pdf_id != ORC_PDF_NONE) { mci.addSamplingFunction(make_pdf(c.pdf_id, c.ndim)); }
    for (int i = 0; i < c.nobs; ++i) {
        mci.addObservable(make_obs(c.obs[i].obs_id, c.ndim), c.obs[i].blocksize, c.obs[i].nskip,
                          c.obs[i].flag_equil != 0, to_estim(c.obs[i].estim_type));
    }
    mci.setTargetAcceptanceRate(c.target_acc);
    mci.setNfindMRT2Iterations(c.nfind);
    mci.setNdecorrelationSteps(c.ndecorr);
}

// Replays the libstdc++ distribution OUTPUTS the main run is going to consume, from a copy of
// MCI's generator taken right before the main run. The reference's own results prove the contract: the C oracle
// consumes exactly this stream and must reproduce avg/acc/x bit for bit.
static int64_t export_draws(const orc_config_t &c, std::mt19937_64 g /*copy*/, double * out, int64_t cap)
{
    std::uniform_real_distribution<double> rsym(-1., 1.), r01(0., 1.);
    const int veclen = std::max(1, c.veclen);
    std::uniform_int_distribution<int> ridx(0, c.ndim/veclen - 1);
    int64_t n = 0;
    auto put = [&](double v) { if (n < cap) { out[n] = v; } ++n; };
    if (c.srrd != ORC_SRRD_UNIFORM) { return -1; } // gaussian draws are stateful in libstdc++ (cached second value): not exported
    for (int64_t s = 0; s < c.nmc; ++s) {
        if (c.pdf_id == ORC_PDF_NONE) { // doStepRandom: ndim x U(0,1)   src/MCIntegrator.cpp:365
            for (int i = 0; i < c.ndim; ++i) { put(r01(g)); }
            continue;
        }
        if (c.move_type == ORC_MOVE_ALL) {
            for (int i = 0; i < c.ndim; ++i) { put(rsym(g)); }
        }
        else if (c.move_type == ORC_MOVE_VEC) {
            put(static_cast<double>(ridx(g)));
            for (int i = 0; i < veclen; ++i) { put(rsym(g)); }
        }
        else {
            const int nsub = c.ms_nsteps > 0 ? c.ms_nsteps : c.ndim;
            for (int k = 0; k < nsub; ++k) {
                put(static_cast<double>(ridx(g)));
                for (int i = 0; i < veclen; ++i) { put(rsym(g)); }
                put(r01(g));
            }
        }
        put(r01(g)); // accept draw, always consumed   src/MCIntegrator.cpp:343
    }
    return n;
}

extern "C" {

const char * mciref_last_error() { return g_err.c_str(); }

// Runs integrate() exactly as a user of the reference would; returns 0 on success.
int mciref_run(const orc_config_t * cfg, orc_result_t * res, orc_trace_t * trace)
{
    try {
        const orc_config_t &c = *cfg;
        MCI mci(c.ndim);
        configure(mci, c);
        const int nobsdim = mci.getNObsDim();
        if (nobsdim > ORC_MAXOBSDIM) { throw std::invalid_argument("ref_harness: nobsdim too large"); }
        std::vector<double> avg(std::max(1, nobsdim), 0.), err(std::max(1, nobsdim), 0.);

        if (trace != nullptr) {
            // split integrate() into its two halves (calibration/decorrelation, then main run): same call sequence
            // as one integrate(nmc, ..., do_find, do_decorr)   src/MCIntegrator.cpp:49-63
            mci.integrate(0, avg.data(), err.data(), c.do_find != 0, c.do_decorr != 0);
            trace->n_draws = export_draws(c, mci._rgen, trace->draws, trace->cap_draws);
            int64_t istep = -1; // first callback comes from initializeSampling()   src/MCIntegrator.cpp:267
            mci.setCallback([&](const MCI &m) {
                if (istep >= 0 && istep < trace->cap_steps) { trace->accepted[istep] = m._wlkstate.accepted ? 1 : 0; }
                ++istep;
            });
            mci.integrate(c.nmc, avg.data(), err.data(), false, false);
            mci.clearCallback();
            trace->n_steps = istep;
        }
        else {
            mci.integrate(c.nmc, avg.data(), err.data(), c.do_find != 0, c.do_decorr != 0);
        }

        std::memset(res, 0, sizeof(*res));
        res->nobsdim = nobsdim;
        std::copy(avg.begin(), avg.begin() + nobsdim, res->avg);
        std::copy(err.begin(), err.begin() + nobsdim, res->err);
        res->acc_rate = mci.getAcceptanceRate();
        for (int i = 0; i < c.ndim; ++i) { res->x_final[i] = mci.getX(i); }
        for (int i = 0; i < std::max(1, c.ntypes); ++i) { res->steps_final[i] = mci.getMRT2Step(i); }
        res->n_acc = mci._acc;
        res->n_rej = mci._rej;
        return 0;
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}

// integrate() with the reference's periodic text dumps switched on (src/MCIntegrator.cpp:495-542): golden files for the
// device-side replacement of storeObservablesOnFile / storeWalkerPositionsOnFile.
int mciref_run_with_files(const orc_config_t * cfg, orc_result_t * res, const char * obs_path, int obs_freq, const char * wlk_path, int wlk_freq)
{
    try {
        const orc_config_t &c = *cfg;
        MCI mci(c.ndim);
        configure(mci, c);
        if (obs_path != nullptr && obs_freq > 0) { mci.storeObservablesOnFile(obs_path, obs_freq); }
        if (wlk_path != nullptr && wlk_freq > 0) { mci.storeWalkerPositionsOnFile(wlk_path, wlk_freq); }
        const int nobsdim = mci.getNObsDim();
        std::vector<double> avg(std::max(1, nobsdim), 0.), err(std::max(1, nobsdim), 0.);
        mci.integrate(c.nmc, avg.data(), err.data(), c.do_find != 0, c.do_decorr != 0);
        std::memset(res, 0, sizeof(*res));
        res->nobsdim = nobsdim;
        std::copy(avg.begin(), avg.begin() + nobsdim, res->avg);
        std::copy(err.begin(), err.begin() + nobsdim, res->err);
        res->acc_rate = mci.getAcceptanceRate();
        return 0;
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}

// integrate() with a step callback (MCI::setCallback, include/mci/MCIntegrator.hpp:186-190) that folds what it sees into four sums:
// golden values for the device-functor replacement.  buf[0] calls, buf[1] accepted calls, buf[2] sum of the first coordinate of
// the state the move leads to, buf[3] sum of the proposed displacement of the last coordinate.
int mciref_run_callback(const orc_config_t * cfg, orc_result_t * res, double * buf)
{
    try {
        const orc_config_t &c = *cfg;
        MCI mci(c.ndim);
        configure(mci, c);
        const int nobsdim = mci.getNObsDim();
        std::vector<double> avg(std::max(1, nobsdim), 0.), err(std::max(1, nobsdim), 0.);
        for (int k = 0; k < 4; ++k) { buf[k] = 0.; }
        const int last = c.ndim - 1;
        mci.setCallback([&](const MCI &m) {
            const WalkerState &w = m._wlkstate;
            buf[0] += 1.;
            buf[1] += w.accepted ? 1. : 0.;
            buf[2] += w.accepted ? w.xnew[0] : w.xold[0];
            buf[3] += w.xnew[last] - w.xold[last];
        });
        mci.integrate(c.nmc, avg.data(), err.data(), c.do_find != 0, c.do_decorr != 0);
        std::memset(res, 0, sizeof(*res));
        res->nobsdim = nobsdim;
        std::copy(avg.begin(), avg.begin() + nobsdim, res->avg);
        std::copy(err.begin(), err.begin() + nobsdim, res->err);
        res->acc_rate = mci.getAcceptanceRate();
        for (int i = 0; i < c.ndim; ++i) { res->x_final[i] = mci.getX(i); }
        return 0;
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}

// Direct access to the reference estimators (include/mci/Estimators.hpp:9-45). nblocks is only used by BLOCK (=100).
int mciref_estimate(int estim_type, int64_t n, int ndim, const double * x, int64_t nblocks, double * avg, double * err)
{
    try {
        switch (estim_type) {
        case ORC_EST_NOOP: NoopEstimator(n, ndim, x, avg, err); break;
        case ORC_EST_UNCORRELATED: UncorrelatedEstimator(n, ndim, x, avg, err); break;
        case ORC_EST_CORRELATED: CorrelatedEstimator(n, ndim, x, avg, err); break;
        case ORC_EST_FCBLOCKER: FCBlockerEstimator(n, ndim, x, avg, err); break;
        case ORC_EST_MJBLOCKER: MJBlockerEstimator(n, ndim, x, avg, err); break;
        case 100:
            if (ndim > 1) { MultiDimBlockEstimator(n, ndim, x, nblocks, avg, err); }
            else { OneDimBlockEstimator(n, x, nblocks, avg[0], err[0]); }
            break;
        default: throw std::invalid_argument("ref_harness: unknown estimator");
        }
        return 0;
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return 1;
    }
}

// TestWalk fixture (test/common/TestMCIFunctions.hpp:14-118): rand()-driven walk used by ut1 and bench_estimators.
// pdf: 0 = SLATER, 1 = GAUSS. Calls srand(seed) first. Returns the acceptance rate.
double mciref_testwalk(int pdf, int nmc, int ndim, double step, double change_prob, unsigned seed, double * datax,
                       uint8_t * datacc, int * nchanged, int * changed_idx)
{
    srand(seed);
    static_assert(sizeof(bool) == 1, "bool size");
    if (pdf == 0) {
        TestWalk<WalkPDF::SLATER> tw(nmc, ndim, step, change_prob);
        tw.generateWalk(datax, reinterpret_cast<bool *>(datacc), nchanged, changed_idx);
        return tw.getAcceptanceRate();
    }
    TestWalk<WalkPDF::GAUSS> tw(nmc, ndim, step, change_prob);
    tw.generateWalk(datax, reinterpret_cast<bool *>(datacc), nchanged, changed_idx);
    return tw.getAcceptanceRate();
}

// Drives one reference accumulator by hand through a recorded walk, the way test/ut1/main.cpp:116-139 does.
// Returns nstore; data_out must hold nstore*nobs doubles (query with data_out == nullptr first).
int64_t mciref_accumulate(int obs_id, int ndim, int blocksize, int nskip, int64_t nmc, const double * datax,
                          const uint8_t * datacc, const int * nchanged, const int * changed_idx, double * data_out)
{
    try {
        auto obs = make_obs(obs_id, ndim);
        auto accu = createAccumulator(*obs, blocksize, nskip);
        accu->allocate(nmc);
        WalkerState wlk(ndim, true);
        for (int64_t i = 0; i < nmc; ++i) {
            std::copy(datax + i*ndim, datax + (i + 1)*ndim, wlk.xnew);
            wlk.accepted = datacc[i] != 0;
            wlk.nchanged = nchanged[i];
            std::copy(changed_idx + i*ndim, changed_idx + i*ndim + wlk.nchanged, wlk.changedIdx);
            accu->accumulate(wlk);
        }
        accu->finalize();
        const int64_t nstore = accu->getNStore();
        if (data_out != nullptr) { std::copy(accu->getData(), accu->getData() + nstore*accu->getNObs(), data_out); }
        return nstore;
    }
    catch (const std::exception &e) {
        g_err = e.what();
        return -1;
    }
}

} // extern "C"
