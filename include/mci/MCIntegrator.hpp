// mci::MCI — the reference's main class (include/mci/MCIntegrator.hpp:26-226) as a facade over the B200 engine's C-ABI.
//
// Same constructor, setters, getters, ownership rules (unique_ptr moves in, const& clones; set*/pop* hand the previous
// object back) and exceptions; integrate() runs the Metropolis walk, acceptance, observables, accumulators and
// estimators on the GPU (libmcig.so). Differences, all additive:
//   * setNWalkers(W): one MCI runs W independent chains ("virtual MPI ranks", src/MPIMCI.cpp:83); W defaults to 1.
//     integrate() returns the MPIMCI combination over walkers (avg of per-walker averages, sqrt(sum err^2)/W).
//   * setRngMode(): Philox4x32-10 in registers (default) or replay of per-walker std::mt19937_64 streams (bit-exact parity).
//   * setCallback takes a device functor (mci/StepCallbackInterface.hpp) instead of a host std::function; dependent observables
//     get their dependencies as a functor argument (mci/DependentObservableInterface.hpp).
//     storeObservablesOnFile / storeWalkerPositionsOnFile work (walker 0, written after the run).
#ifndef MCIG_MCI_MCINTEGRATOR_HPP
#define MCIG_MCI_MCINTEGRATOR_HPP

#include "mci/DependentObservableInterface.hpp"
#include "mci/DomainInterface.hpp"
#include "mci/Factories.hpp"
#include "mci/ObservableFunctionInterface.hpp"
#include "mci/SamplingFunctionInterface.hpp"
#include "mci/StepCallbackInterface.hpp"
#include "mci/TrialMoveInterface.hpp"

#include <cstdint>
#include <functional>
#include <iostream>
#include <memory>
#include <random>
#include <string>
#include <vector>

namespace mci
{
enum class RngMode { Philox32 = MCIG_RNG_PHILOX32, Philox53 = MCIG_RNG_PHILOX53, Replay = MCIG_RNG_REPLAY };

class MCI
{
private:
    const int _ndim;
    mcig_ctx * _ctx;
    std::unique_ptr<DomainInterface> _domain;
    std::unique_ptr<TrialMoveInterface> _trialMove;
    std::vector<std::unique_ptr<SamplingFunctionInterface>> _pdfs;
    struct ObsElement
    {
        std::unique_ptr<ObservableFunctionInterface> obs;
        int blocksize, nskip;
        bool flag_equil;
        EstimatorType estimType;
    };
    std::vector<ObsElement> _obs;
    int _nobsdim{0};
    int _NfindMRT2Iterations{-50};      // src/MCIntegrator.cpp:638
    int64_t _NdecorrelationSteps{-10000}; // :639
    double _targetaccrate{0.5};         // :637
    bool _dirtyMove{true}, _dirtyPdf{true}, _dirtyObs{true}, _dirtyDomain{true}, _dirtyCback{false};
    std::unique_ptr<StepCallbackInterface> _cback; // device-functor replacement of the reference's std::function _cback
    int64_t _cbackDoubles{0};
    std::mt19937_64 _hostgen; // only for newRandomX()/moveX() (manual position helpers)
    mutable std::vector<double> _xcache; // backing store of getX() const

    static int toSrrd(SRRDType t)
    {
        return static_cast<int>(t); // same enumerator order as MCIG_SRRD_* (include/mci/Factories.hpp:119-133)
    }

    void pushConfiguration()
    {
        pushDomainOnly();
        if (_dirtyMove) {
            const TrialMoveInterface & m = *_trialMove;
            int mt = MCIG_MOVE_ALL;
            if (m.getMoveType() == MoveType::Vec) { mt = MCIG_MOVE_VEC; }
            if (m.getMoveType() == MoveType::MultiStep) { mt = MCIG_MOVE_MULTISTEP; }
            const DeviceFunctor mf = m.deviceFunctor();
            if (!mf.name.empty()) { // user-defined move: its device functor
                detail::check(mcig_set_move_plugin(_ctx, mf.resolve(MCIG_PLUGIN_MOVE, _ndim, m.getNUniforms()), mf.params.data(), static_cast<int>(mf.params.size()),
                                                   m.getNTypes(), m.getNTypes() > 1 ? m.getTypeEnds() : nullptr));
            }
            else {
                detail::check(mcig_set_move(_ctx, mt, toSrrd(m.getSRRDType()), m.getVecLen(), m.getNTypes(), m.getNTypes() > 1 ? m.getTypeEnds() : nullptr));
            }
            if (mf.name.empty() && !m.getSRRDParams().empty()) { detail::check(mcig_set_srrd_params(_ctx, static_cast<int>(m.getSRRDParams().size()), m.getSRRDParams().data())); }
            if (mt == MCIG_MOVE_MULTISTEP) {
                const auto & ms = dynamic_cast<const MultiStepMove &>(m);
                detail::check(mcig_multistep_config(_ctx, ms.getNSteps()));
                for (int i = 0; i < ms.getNPDF(); ++i) {
                    const SamplingFunctionInterface & pdf = ms.getSamplingFunction(i);
                    const DeviceFunctor f = pdf.deviceFunctor();
                    detail::check(mcig_multistep_add_pdf(_ctx, f.resolve(MCIG_PLUGIN_PDF, pdf.getNDim(), pdf.getNProto()), f.params.data(), static_cast<int>(f.params.size())));
                }
            }
            _dirtyMove = false;
        }
        for (int i = 0; i < _trialMove->getNStepSizes(); ++i) { detail::check(mcig_set_step(_ctx, i, _trialMove->getStepSize(i))); }
        if (_dirtyPdf) {
            detail::check(mcig_clear_pdfs(_ctx));
            for (auto & pdf : _pdfs) {
                const DeviceFunctor f = pdf->deviceFunctor();
                detail::check(mcig_add_pdf(_ctx, f.resolve(MCIG_PLUGIN_PDF, pdf->getNDim(), pdf->getNProto()), f.params.data(), static_cast<int>(f.params.size())));
            }
            _dirtyPdf = false;
        }
        if (_dirtyObs) {
            detail::check(mcig_clear_obs(_ctx));
            for (auto & el : _obs) {
                DeviceFunctor f = el.obs->deviceFunctor();
                if (dynamic_cast<const DependentObservableInterface *>(el.obs.get()) != nullptr) { f.dependent = true; }
                detail::check(mcig_add_obs(_ctx, f.resolve(MCIG_PLUGIN_OBS, el.obs->getNDim(), el.obs->getNObs()), f.params.data(), static_cast<int>(f.params.size()),
                                           el.blocksize, el.nskip, el.flag_equil ? 1 : 0, static_cast<int>(el.estimType)));
            }
            _dirtyObs = false;
        }
        if (_dirtyCback) {
            if (_cback) {
                const DeviceFunctor f = _cback->deviceFunctor();
                _cbackDoubles = _cback->bufferDoubles(mcig_get_walkers(_ctx));
                detail::check(mcig_set_callback(_ctx, f.resolve(MCIG_PLUGIN_CALLBACK, 0, 0), f.params.data(), static_cast<int>(f.params.size()), _cbackDoubles));
            }
            else {
                _cbackDoubles = 0;
                detail::check(mcig_clear_callback(_ctx));
            }
            _dirtyCback = false;
        }
        detail::check(mcig_set_autotune(_ctx, _NfindMRT2Iterations, _NdecorrelationSteps, _targetaccrate));
    }

public:
    explicit MCI(int ndim): _ndim(ndim), _ctx(mcig_create(ndim))
    {
        if (_ctx == nullptr) { throw std::invalid_argument(mcig_last_error()); }
        _domain.reset(new UnboundDomain(_ndim));      // default: unbound domain (src/MCIntegrator.cpp:631)
        _trialMove = createMoveDefault(MoveType::All, _ndim); // default: uniform all-move (:634)
        std::random_device rdev;
        _hostgen.seed(rdev());
    }
    ~MCI() { mcig_destroy(_ctx); }
    MCI(const MCI &) = delete;
    MCI & operator=(const MCI &) = delete;

    // --- Setters
    void setSeed(uint_fast64_t seed)
    {
        detail::check(mcig_set_seed(_ctx, seed));
        _hostgen.seed(seed);
    }
    void setX(int i, double val)
    {
        std::vector<double> x(static_cast<size_t>(_ndim));
        detail::check(mcig_get_x(_ctx, 0, x.data()));
        x[static_cast<size_t>(i)] = val;
        setX(x.data());
    }
    void setX(const double x[])
    {
        pushDomainOnly();
        if (!_domain->deviceFunctor().name.empty()) { // user-defined domain: applied here, by its host twin (src/MCIntegrator.cpp:600-604)
            std::vector<double> t(x, x + _ndim);
            _domain->applyDomain(t.data());
            detail::check(mcig_set_x(_ctx, t.data()));
            return;
        }
        detail::check(mcig_set_x(_ctx, x));
    }
    void moveX()
    { // single manual move with the configured step size (uniform), then the domain (src/MCIntegrator.cpp:606-612)
        std::vector<double> x(static_cast<size_t>(_ndim));
        detail::check(mcig_get_x(_ctx, 0, x.data()));
        std::uniform_real_distribution<double> rd(-1., 1.);
        for (int i = 0; i < _ndim; ++i) { x[static_cast<size_t>(i)] += _trialMove->getStepSize(_trialMove->getStepSizeIndex(i))*rd(_hostgen); }
        setX(x.data());
    }
    void newRandomX()
    { // src/MCIntegrator.cpp:614-619
        std::vector<double> x(static_cast<size_t>(_ndim));
        std::uniform_real_distribution<double> rd(0., 1.);
        for (auto & v : x) { v = rd(_hostgen); }
        _domain->scaleToDomain(x.data());
        setX(x.data());
    }
    void centerX()
    {
        std::vector<double> x(static_cast<size_t>(_ndim));
        _domain->getCenter(x.data());
        setX(x.data());
    }

    void setMRT2Step(double mrt2step)
    {
        for (int i = 0; i < _trialMove->getNStepSizes(); ++i) { _trialMove->setStepSize(i, mrt2step); }
    }
    void setMRT2Step(int i, double mrt2step)
    {
        if (i < _trialMove->getNStepSizes()) { _trialMove->setStepSize(i, mrt2step); }
        else { std::cout << "[MCI::setMRT2Step] Warning: Tried to set non-existing MRT2step index." << std::endl; }
    }
    void setMRT2Step(const double mrt2step[])
    {
        for (int i = 0; i < _trialMove->getNStepSizes(); ++i) { _trialMove->setStepSize(i, mrt2step[i]); }
    }

    void setTargetAcceptanceRate(double targetaccrate) { _targetaccrate = targetaccrate; }
    void setNfindMRT2Iterations(int niterations) { _NfindMRT2Iterations = niterations; }
    void setNdecorrelationSteps(int64_t nsteps) { _NdecorrelationSteps = nsteps; }

    // --- Domain
    std::unique_ptr<DomainInterface> setDomain(std::unique_ptr<DomainInterface> domain)
    {
        if (domain->ndim != _ndim) { throw std::invalid_argument("[MCI::setDomain] Passed domain's number of dimensions is not equal to MCI's number of walkers."); }
        std::swap(domain, _domain);
        _dirtyDomain = true;
        pushDomainOnly(); // applies the new domain to the current position, as the reference does
        return domain;
    }
    std::unique_ptr<DomainInterface> setDomain(const DomainInterface & domain) { return this->setDomain(domain.clone()); }
    std::unique_ptr<DomainInterface> resetDomain() { return this->setDomain(std::unique_ptr<DomainInterface>(new UnboundDomain(_ndim))); }
    void setIRange(double lbound, double ubound) { this->setDomain(std::unique_ptr<DomainInterface>(new OrthoPeriodicDomain(_ndim, lbound, ubound))); }
    void setIRange(const double lbounds[], const double ubounds[]) { this->setDomain(std::unique_ptr<DomainInterface>(new OrthoPeriodicDomain(_ndim, lbounds, ubounds))); }

    // --- Trial moves
    std::unique_ptr<TrialMoveInterface> setTrialMove(std::unique_ptr<TrialMoveInterface> tmove)
    {
        if (tmove->getNDim() != _ndim) { throw std::invalid_argument("[MCI::setTrialMove] Passed trial move's number of inputs is not equal to MCI's number of walkers."); }
        toSrrd(tmove->getSRRDType());
        std::swap(tmove, _trialMove);
        _dirtyMove = true;
        return tmove;
    }
    std::unique_ptr<TrialMoveInterface> setTrialMove(const TrialMoveInterface & tmove) { return this->setTrialMove(tmove.clone()); }
    std::unique_ptr<TrialMoveInterface> setTrialMove(MoveType move) { return this->setTrialMove(createMoveDefault(move, _ndim)); }
    std::unique_ptr<TrialMoveInterface> setTrialMove(SRRDType srrd, int veclen = 0, int ntypes = 1, int typeEnds[] = nullptr)
    {
        if (veclen > 0) {
            if (_ndim%veclen != 0) { throw std::invalid_argument("[MCI::setTrialMove] MCI's number of walkers must be a multiple of passed veclen."); }
            return this->setTrialMove(createSRRDVecMove(srrd, _ndim/veclen, veclen, ntypes, typeEnds));
        }
        return this->setTrialMove(createSRRDAllMove(srrd, _ndim, ntypes, typeEnds));
    }

    // --- Observables
    void addObservable(std::unique_ptr<ObservableFunctionInterface> obs, int blocksize, int nskip, bool flag_equil, EstimatorType estimType)
    {
        blocksize = std::max(0, blocksize);
        nskip = std::max(1, nskip);
        if (obs->getNDim() != _ndim) { throw std::invalid_argument("[MCI::addObservable] Passed observable function's number of inputs is not equal to MCI's number of walkers."); }
        if (flag_equil && estimType == EstimatorType::Noop) {
            throw std::invalid_argument("[MCI::addObservable] Requested automatic observable equilibration requires estimator with error calculation.");
        }
        _nobsdim += obs->getNObs();
        _obs.push_back(ObsElement{std::move(obs), blocksize, nskip, flag_equil, estimType});
        _dirtyObs = true;
    }
    void addObservable(const ObservableFunctionInterface & obs, int blocksize, int nskip, bool flag_equil, EstimatorType estimType)
    {
        this->addObservable(obs.clone(), blocksize, nskip, flag_equil, estimType);
    }
    void addObservable(std::unique_ptr<ObservableFunctionInterface> obs, int blocksize, int nskip, bool flag_equil, bool flag_correlated)
    {
        this->addObservable(std::move(obs), blocksize, nskip, flag_equil, selectEstimatorType(flag_correlated, blocksize > 0));
    }
    void addObservable(const ObservableFunctionInterface & obs, int blocksize, int nskip, bool flag_equil, bool flag_correlated)
    {
        this->addObservable(obs.clone(), blocksize, nskip, flag_equil, flag_correlated);
    }
    void addObservable(std::unique_ptr<ObservableFunctionInterface> obs, int blocksize = 1, int nskip = 1)
    {
        this->addObservable(std::move(obs), blocksize, nskip, blocksize > 0, blocksize == 1);
    }
    void addObservable(const ObservableFunctionInterface & obs, int blocksize = 1, int nskip = 1) { this->addObservable(obs.clone(), blocksize, nskip); }
    std::unique_ptr<ObservableFunctionInterface> popObservable()
    {
        auto obs = std::move(_obs.back().obs);
        _obs.pop_back();
        _nobsdim -= obs->getNObs();
        _dirtyObs = true;
        return obs;
    }
    void clearObservables()
    {
        _obs.clear();
        _nobsdim = 0;
        _dirtyObs = true;
    }

    // --- Sampling functions
    void addSamplingFunction(std::unique_ptr<SamplingFunctionInterface> pdf)
    {
        if (pdf->getNDim() != _ndim) { throw std::invalid_argument("[MCI::addSamplingFunction] Passed sampling function's number of inputs is not equal to MCI's number of walkers."); }
        _pdfs.emplace_back(std::move(pdf));
        _dirtyPdf = true;
    }
    void addSamplingFunction(const SamplingFunctionInterface & pdf) { this->addSamplingFunction(pdf.clone()); }
    std::unique_ptr<SamplingFunctionInterface> popSamplingFunction()
    {
        auto pdf = std::move(_pdfs.back());
        _pdfs.pop_back();
        _dirtyPdf = true;
        return pdf;
    }
    void clearSamplingFunctions()
    {
        _pdfs.clear();
        _dirtyPdf = true;
    }

    // --- Callback: a device functor called after every move (see mci/StepCallbackInterface.hpp). A host std::function cannot run
    // inside a kernel: that overload throws and names the replacement.
    void setCallback(const StepCallbackInterface & cback)
    {
        _cback = cback.clone();
        _dirtyCback = true;
    }
    void setCallback(const std::function<void(const MCI &)> &)
    {
        throw std::logic_error("[MCI::setCallback] a host callback cannot run inside the device-resident walk: pass a mci::StepCallbackInterface (device functor)");
    }
    void clearCallback()
    {
        _cback.reset();
        _dirtyCback = true;
    }
    std::vector<double> getCallbackBuffer() const
    {
        std::vector<double> out(static_cast<size_t>(_cbackDoubles));
        detail::check(mcig_get_callback_buffer(_ctx, out.data(), _cbackDoubles));
        return out;
    }
    // file dumps of walker 0 every freq-th step (src/MCIntegrator.cpp:495-542), written after the run from device-side accumulators
    void storeObservablesOnFile(const std::string & filepath, int freq) { detail::check(mcig_store_on_file(_ctx, 0, filepath.c_str(), freq)); }
    void clearObservableFile() { detail::check(mcig_store_on_file(_ctx, 0, "", 0)); }
    void storeWalkerPositionsOnFile(const std::string & filepath, int freq) { detail::check(mcig_store_on_file(_ctx, 1, filepath.c_str(), freq)); }
    void clearWalkerFile() { detail::check(mcig_store_on_file(_ctx, 1, "", 0)); }

    // --- Getters
    int getNDim() const { return _ndim; }
    double getX(int i) const
    {
        std::vector<double> x(static_cast<size_t>(_ndim));
        detail::check(mcig_get_x(_ctx, 0, x.data()));
        return x[static_cast<size_t>(i)];
    }
    const double * getX() const // walker 0, as the reference's pointer to its xold (valid until the next call on this object)
    {
        _xcache.resize(static_cast<size_t>(_ndim));
        detail::check(mcig_get_x(_ctx, 0, _xcache.data()));
        return _xcache.data();
    }
    void getX(double x[], int64_t walker = 0) const { detail::check(mcig_get_x(_ctx, walker, x)); }
    double getMRT2Step(int i) const { return (i < _trialMove->getNStepSizes()) ? _trialMove->getStepSize(i) : 0.; }
    double getTargetAcceptanceRate() const { return _targetaccrate; }
    double getAcceptanceRate() const { return mcig_get_acceptance_rate(_ctx); }
    int getNfindMRT2Iterations() const { return _NfindMRT2Iterations; }
    int64_t getNdecorrelationSteps() const { return _NdecorrelationSteps; }
    const DomainInterface & getDomain() const { return *_domain; }
    TrialMoveInterface & getTrialMove() const { return *_trialMove; }
    SamplingFunctionInterface & getSamplingFunction(int i) const { return *_pdfs[static_cast<size_t>(i)]; }
    int getNPDF() const { return static_cast<int>(_pdfs.size()); }
    ObservableFunctionInterface & getObservable(int i) const { return *_obs[static_cast<size_t>(i)].obs; }
    int getNObs() const { return static_cast<int>(_obs.size()); }
    int getNObsDim() const { return _nobsdim; }

    // --- Integrate (src/MCIntegrator.cpp:43-82)
    void integrate(int64_t Nmc, double average[], double error[], bool doFindMRT2step = true, bool doDecorrelation = true)
    {
        pushConfiguration();
        const int rc = mcig_integrate(_ctx, Nmc, average, error, doFindMRT2step ? 1 : 0, doDecorrelation ? 1 : 0);
        for (int i = 0; i < _trialMove->getNStepSizes(); ++i) { _trialMove->setStepSize(i, mcig_get_step(_ctx, i)); } // calibrated sizes
        detail::check(rc);
    }

    // --- Engine extensions
    void setNWalkers(int64_t nwalkers, int64_t globalOffset = 0, int64_t totalWalkers = -1)
    {
        detail::check(mcig_set_walkers(_ctx, nwalkers, globalOffset, totalWalkers < 0 ? nwalkers + globalOffset : totalWalkers));
        _dirtyCback = (_cback != nullptr) || _dirtyCback; // its buffer is sized by the walker count
    }
    int64_t getNWalkers() const { return mcig_get_walkers(_ctx); }
    void setRngMode(RngMode mode) { detail::check(mcig_set_rng_mode(_ctx, static_cast<int>(mode))); }
    void setWalkerSeeds(const uint64_t seeds[], int64_t n) { detail::check(mcig_set_walker_seeds(_ctx, seeds, n)); }
    void setDevice(int device) { detail::check(mcig_set_device(_ctx, device)); }
    void setPhiloxRounds(int rounds) { detail::check(mcig_set_philox_rounds(_ctx, rounds)); }
    void setAllreduce(mcig_allreduce_fn fn, void * user) { detail::check(mcig_set_allreduce(_ctx, fn, user)); }
    void getCrossWalkerError(double err[]) const { detail::check(mcig_get_cross_walker_error(_ctx, err, mcig_get_result_nobsdim(_ctx))); }
    void setKeepSamples(bool on) { detail::check(mcig_set_keep_samples(_ctx, on ? 1 : 0)); }
    // engine extras without reference analogue: compile + load every kernel of the configuration now (prebuild), and additionally rehearse the coming
    // integrate call with every trace undone, so that the first real call runs at steady-state speed (warmup; include/mcig.h)
    void prebuild() { pushConfiguration(); detail::check(mcig_prebuild(_ctx)); }
    void warmup(int64_t Nmc, bool doFindMRT2step = true, bool doDecorrelation = true)
    {
        pushConfiguration();
        detail::check(mcig_warmup(_ctx, Nmc, doFindMRT2step ? 1 : 0, doDecorrelation ? 1 : 0));
    }
    void attachComm(bool on = true) { detail::check(mcig_attach_comm(_ctx, on ? 1 : 0)); }
    void getWalkerResults(double avg[], double err[]) const { detail::check(mcig_get_walker_results(_ctx, avg, err)); }
    mcig_ctx * handle() const { return _ctx; }

private:
    void pushDomainOnly()
    {
        if (!_dirtyDomain) { return; }
        const DeviceFunctor df = _domain->deviceFunctor();
        if (!df.name.empty()) { // a user-defined domain: its device functor, plus getSizes / getVolume for the host-side rules
            std::vector<double> sizes(static_cast<size_t>(_ndim)), x(static_cast<size_t>(_ndim));
            _domain->getSizes(sizes.data());
            detail::check(mcig_get_x(_ctx, 0, x.data()));
            detail::check(mcig_set_domain_plugin(_ctx, df.resolve(MCIG_PLUGIN_DOMAIN, _ndim, 0), df.params.data(), static_cast<int>(df.params.size()), sizes.data(), _domain->getVolume()));
            _domain->applyDomain(x.data()); // (the engine applies its built-in domains to the position itself)
            detail::check(mcig_set_x(_ctx, x.data()));
        }
        else if (_domain->isPeriodic()) {
            std::vector<double> lb(static_cast<size_t>(_ndim)), ub(static_cast<size_t>(_ndim));
            _domain->getBounds(lb.data(), ub.data());
            detail::check(mcig_set_domain_ortho(_ctx, lb.data(), ub.data()));
        }
        else { detail::check(mcig_set_domain_unbound(_ctx)); }
        _dirtyDomain = false;
    }
};
} // namespace mci
#endif
