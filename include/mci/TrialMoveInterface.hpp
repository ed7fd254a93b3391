// Trial moves — mirror of include/mci/TrialMoveInterface.hpp:16-70, TypedMoveInterface.hpp:20-63, SRRDAllMove.hpp,
// SRRDVecMove.hpp, MultiStepMove.hpp. The proposal itself runs in the walk kernel (device/mcig_device.cuh, MOVE 0/1/2);
// these host objects carry the move's configuration (kind, vector length, typed step sizes, MultiStep sub-move and
// sub-sampling functions) and the step-size accessors findMRT2Step needs. All ten SRRDType distributions have device samplers
// (device/mcig_device.cuh: Proposal); MultiStepMove sub-moves are uniform single-vector moves.
#ifndef MCIG_MCI_TRIALMOVEINTERFACE_HPP
#define MCIG_MCI_TRIALMOVEINTERFACE_HPP

#include "mci/Clonable.hpp"
#include "mci/DeviceFunctor.hpp"
#include "mci/SamplingFunctionInterface.hpp"

#include <algorithm>
#include <memory>
#include <random>
#include <stdexcept>
#include <string>
#include <vector>

namespace mci
{
enum class MoveType { All, Vec, MultiStep };                  // include/mci/Factories.hpp:108-114
enum class SRRDType { Uniform, Gaussian, Student, Cauchy, Exponential, Gamma, Weibull, Lognormal, Chisq, Fisher }; // :119-133
static constexpr double DEFAULT_MRT2STEP = 0.05;              // include/mci/Factories.hpp:145

class TrialMoveInterface: public Clonable<TrialMoveInterface>
{
protected:
    const int _ndim;
    explicit TrialMoveInterface(int ndim): _ndim(ndim) {}

public:
    int getNDim() const { return _ndim; }
    bool hasStepSizes() const { return this->getNStepSizes() > 0; }
    virtual MoveType getMoveType() const = 0;
    virtual SRRDType getSRRDType() const { return SRRDType::Uniform; }
    virtual const std::vector<double> & getSRRDParams() const // parameters of a pre-made distribution passed to the constructor; empty = defaults
    {
        static const std::vector<double> none;
        return none;
    }
    // A user-defined move (the reference's interface is meant to be subclassed: include/mci/TrialMoveInterface.hpp:16-70, trialMove(WalkerState &, ...)):
    // return the device twin of trialMove here (plugin kind MCIG_PLUGIN_MOVE, contract at mcig_set_move_plugin in include/mcig.h) and the number of
    // uniforms in [0,1) it consumes per step (0 = one per coordinate); getMoveType() is MoveType::All for such a move (all coordinates may change)
    virtual DeviceFunctor deviceFunctor() const { return DeviceFunctor(); }
    virtual int getNUniforms() const { return 0; }
    virtual int getVecLen() const { return 1; }
    virtual int getNTypes() const = 0;
    virtual const int * getTypeEnds() const = 0;
    virtual int getNStepSizes() const = 0;
    virtual double getStepSize(int i) const = 0;
    virtual void setStepSize(int i, double val) = 0;
    virtual double getChangeRate() const = 0;
    virtual int getStepSizeIndex(int xidx) const = 0;
    void scaleStepSize(int i, double fac) { this->setStepSize(i, this->getStepSize(i)*fac); }
    void scaleStepSizes(double fac)
    {
        for (int i = 0; i < this->getNStepSizes(); ++i) { this->scaleStepSize(i, fac); }
    }
};

// typed step sizes: x = (a0,a1,a2,b0,b1,c0) -> ntypes 3, typeEnds (3,5,6)
class TypedMoveInterface: public TrialMoveInterface
{
protected:
    const int _ntypes;
    std::vector<int> _typeEnds;
    std::vector<double> _stepSizes;
    const SRRDType _srrd;
    std::vector<double> _srrdPar; // see mcig_set_srrd_params (include/mcig.h)
    TypedMoveInterface(int ndim, int ntypes, const int typeEnds[], double initStepSize, SRRDType srrd):
            TrialMoveInterface(ndim), _ntypes(ntypes), _srrd(srrd)
    {
        if (_ntypes < 1) { throw std::invalid_argument("[TypedMoveInterface] Number of types must be at least 1."); }
        if (_ntypes > 1) {
            if (typeEnds == nullptr) { throw std::invalid_argument("[TypedMoveInterface] When ntypes>1, passed typeEnds must not be null."); }
            _typeEnds.assign(typeEnds, typeEnds + _ntypes);
        }
        else { _typeEnds.assign(1, ndim); }
        _stepSizes.assign(static_cast<size_t>(_ntypes), initStepSize);
    }

public:
    SRRDType getSRRDType() const final { return _srrd; }
    const std::vector<double> & getSRRDParams() const final { return _srrdPar; }
    int getNTypes() const final { return _ntypes; }
    const int * getTypeEnds() const final { return _typeEnds.data(); }
    int getNStepSizes() const final { return _ntypes; }
    void setStepSize(int i, double val) final { _stepSizes[static_cast<size_t>(i)] = val; }
    double getStepSize(int i) const final { return _stepSizes[static_cast<size_t>(i)]; }
    int getStepSizeIndex(int xidx) const final
    {
        for (int i = 0; i < _ntypes; ++i) {
            if (xidx < _typeEnds[static_cast<size_t>(i)]) { return i; }
        }
        throw std::runtime_error("[TypedMoveInterface::getUsedStepSize] Passed xidx exceeds expected range.");
    }
};

// Symmetrised positive distribution (include/mci/TrialMoveInterface.hpp:81-98): value from prrd, sign from a fair coin. Kept so that user code
// building SymmetrizedPRRD<std::gamma_distribution<double>> objects for a move constructor compiles unchanged; the device engine reads its parameters.
template <class PRRD>
struct SymmetrizedPRRD
{
private:
    std::bernoulli_distribution bd;

public:
    PRRD prrd;
    SymmetrizedPRRD() = default;
    explicit SymmetrizedPRRD(PRRD myPRRD) { prrd = myPRRD; }
    double operator()(std::mt19937_64 &rgen)
    {
        const double val = prrd(rgen);
        return (bd(rgen)) ? val : -val;
    }
};

// The reference's move types are templates over the distribution TYPE and take an optional pre-made distribution OBJECT (`const SRRD * rdist`,
// include/mci/SRRDAllMove.hpp:45-58). Here the distribution is a tag plus parameters handed to the kernel generator: SRRDDist maps each tag to the
// reference's distribution type and reads the parameters out of such an object. Locations must be 0 (a move distribution has to be symmetric).
namespace detail
{
inline void requireSymmetric(bool ok, const char * what)
{
    if (!ok) { throw std::invalid_argument(std::string("[mcig] ") + what + ": the device engine takes move distributions that are symmetric around 0"); }
}
inline void checkUniformRdist(const std::uniform_real_distribution<double> * rdist)
{
    if (rdist != nullptr && !(rdist->a() == -1. && rdist->b() == 1.)) {
        throw std::invalid_argument("[mcig] uniform move distributions other than (-1, 1) are expressed through the step size");
    }
}
} // namespace detail
template <SRRDType T> struct SRRDDist;
template <> struct SRRDDist<SRRDType::Gaussian> {
    typedef std::normal_distribution<double> type;
    static std::vector<double> params(const type &d) { detail::requireSymmetric(d.mean() == 0., "normal_distribution with a mean"); return {d.stddev()}; }
};
template <> struct SRRDDist<SRRDType::Student> {
    typedef std::student_t_distribution<double> type;
    static std::vector<double> params(const type &d) { return {d.n()}; }
};
template <> struct SRRDDist<SRRDType::Cauchy> {
    typedef std::cauchy_distribution<double> type;
    static std::vector<double> params(const type &d) { detail::requireSymmetric(d.a() == 0., "cauchy_distribution with a location"); return {d.b()}; }
};
template <> struct SRRDDist<SRRDType::Exponential> {
    typedef SymmetrizedPRRD<std::exponential_distribution<double>> type;
    static std::vector<double> params(const type &d) { return {d.prrd.lambda()}; }
};
template <> struct SRRDDist<SRRDType::Gamma> {
    typedef SymmetrizedPRRD<std::gamma_distribution<double>> type;
    static std::vector<double> params(const type &d) { return {d.prrd.alpha(), d.prrd.beta()}; }
};
template <> struct SRRDDist<SRRDType::Weibull> {
    typedef SymmetrizedPRRD<std::weibull_distribution<double>> type;
    static std::vector<double> params(const type &d) { return {d.prrd.a(), d.prrd.b()}; }
};
template <> struct SRRDDist<SRRDType::Lognormal> {
    typedef SymmetrizedPRRD<std::lognormal_distribution<double>> type;
    static std::vector<double> params(const type &d) { return {d.prrd.m(), d.prrd.s()}; }
};
template <> struct SRRDDist<SRRDType::Chisq> {
    typedef SymmetrizedPRRD<std::chi_squared_distribution<double>> type;
    static std::vector<double> params(const type &d) { return {d.prrd.n()}; }
};
template <> struct SRRDDist<SRRDType::Fisher> {
    typedef SymmetrizedPRRD<std::fisher_f_distribution<double>> type;
    static std::vector<double> params(const type &d) { return {d.prrd.m(), d.prrd.n()}; }
};

// all-particle move with a symmetric real-valued random distribution (include/mci/SRRDAllMove.hpp)
class SRRDAllMove: public TypedMoveInterface
{
    TrialMoveInterface * _clone() const final { return new SRRDAllMove(*this); }

protected:
    void setSRRDParams(std::vector<double> par) { _srrdPar = std::move(par); }

public:
    // UniformAllMove(ndim, step, &uniform_real_distribution(-1, 1)): the reference's uniform instantiation with its optional pre-made distribution
    SRRDAllMove(int ndim, int ntypes, const int typeEnds[], double initStepSize, const std::uniform_real_distribution<double> * rdist):
            SRRDAllMove(ndim, ntypes, typeEnds, initStepSize, SRRDType::Uniform) { detail::checkUniformRdist(rdist); }
    SRRDAllMove(int ndim, double initStepSize, const std::uniform_real_distribution<double> * rdist): SRRDAllMove(ndim, 1, nullptr, initStepSize, rdist) {}
    SRRDAllMove(int ndim, int ntypes, const int typeEnds[], double initStepSize, SRRDType srrd = SRRDType::Uniform):
            TypedMoveInterface(ndim, ntypes, typeEnds, initStepSize, srrd) {}
    SRRDAllMove(int ndim, double initStepSize, SRRDType srrd = SRRDType::Uniform): SRRDAllMove(ndim, 1, nullptr, initStepSize, srrd) {}
    MoveType getMoveType() const final { return MoveType::All; }
    double getChangeRate() const final { return 1.; }
};

// single-vector move (include/mci/SRRDVecMove.hpp)
class SRRDVecMove: public TypedMoveInterface
{
    const int _nvecs, _veclen;
    // The reference's vec-move clone does not pass its distribution on (include/mci/SRRDVecMove.hpp:30-33: `new SRRDVecMove(_nvecs, _veclen, _ntypes,
    // _typeEnds, _stepSizes)`, unlike SRRDAllMove.hpp:34-37), and MCI::setTrialMove(const TrialMoveInterface &) stores a clone: a pre-made distribution
    // handed to a VEC move never reaches the sampling loop there -- the default-constructed one is used. Mirrored, so that the same program gives the
    // same numbers (pinned by the par_*_vec goldens, tests/configs.py); mcig_set_srrd_params itself works for vec-moves too.
    TrialMoveInterface * _clone() const final
    {
        SRRDVecMove * c = new SRRDVecMove(*this);
        c->_srrdPar.clear();
        return c;
    }

protected:
    void setSRRDParams(std::vector<double> par) { _srrdPar = std::move(par); }

public:
    SRRDVecMove(int nvecs, int veclen, int ntypes, const int typeEnds[], double initStepSize, const std::uniform_real_distribution<double> * rdist):
            SRRDVecMove(nvecs, veclen, ntypes, typeEnds, initStepSize, SRRDType::Uniform) { detail::checkUniformRdist(rdist); }
    SRRDVecMove(int nvecs, int veclen, double initStepSize, const std::uniform_real_distribution<double> * rdist):
            SRRDVecMove(nvecs, veclen, 1, nullptr, initStepSize, rdist) {}
    SRRDVecMove(int nvecs, int veclen, int ntypes, const int typeEnds[], double initStepSize, SRRDType srrd = SRRDType::Uniform):
            TypedMoveInterface(nvecs*veclen, ntypes, typeEnds, initStepSize, srrd), _nvecs(nvecs), _veclen(veclen)
    {
        if (_nvecs < 1) { throw std::invalid_argument("[SRRDVecMove] Number of vectors must be at least 1."); }
        if (_veclen < 1) { throw std::invalid_argument("[SRRDVecMove] Vector length must be at least 1."); }
        if (_ntypes > 1) {
            for (int e : _typeEnds) {
                if (e%_veclen != 0) { throw std::invalid_argument("[SRRDVecMove] All type end indices must be multiples of vector length."); }
            }
        }
    }
    SRRDVecMove(int nvecs, int veclen, double initStepSize, SRRDType srrd = SRRDType::Uniform): SRRDVecMove(nvecs, veclen, 1, nullptr, initStepSize, srrd) {}
    MoveType getMoveType() const final { return MoveType::Vec; }
    int getVecLen() const final { return _veclen; }
    double getChangeRate() const final { return 1./_nvecs; }
};

// names of the reference's uniform instantiations (include/mci/SRRDAllMove.hpp:84, SRRDVecMove.hpp:100): the distribution is a
// run-time tag here (default uniform), so UniformAllMove(ndim, step) / UniformVecMove(nvecs, veclen, step) construct as in the reference
using UniformAllMove = SRRDAllMove;
using UniformVecMove = SRRDVecMove;

// the other named instantiations (include/mci/SRRDAllMove.hpp:85-95, SRRDVecMove.hpp:101-111): same constructors, distribution fixed by the type
// and the optional pre-made distribution (e.g. StudentAllMove(ndim, 0.05, &student_t(2)), test/ut5/main.cpp:110-113)
template <SRRDType T>
struct SRRDAllMoveOf final: public SRRDAllMove
{
    typedef typename SRRDDist<T>::type Dist;
    SRRDAllMoveOf(int ndim, double initStepSize, const Dist * rdist = nullptr): SRRDAllMoveOf(ndim, 1, nullptr, initStepSize, rdist) {}
    SRRDAllMoveOf(int ndim, int ntypes, const int typeEnds[], double initStepSize, const Dist * rdist = nullptr): SRRDAllMove(ndim, ntypes, typeEnds, initStepSize, T)
    {
        if (rdist != nullptr) { this->setSRRDParams(SRRDDist<T>::params(*rdist)); }
    }
};
template <SRRDType T>
struct SRRDVecMoveOf final: public SRRDVecMove
{
    typedef typename SRRDDist<T>::type Dist;
    SRRDVecMoveOf(int nvecs, int veclen, double initStepSize, const Dist * rdist = nullptr): SRRDVecMoveOf(nvecs, veclen, 1, nullptr, initStepSize, rdist) {}
    SRRDVecMoveOf(int nvecs, int veclen, int ntypes, const int typeEnds[], double initStepSize, const Dist * rdist = nullptr):
            SRRDVecMove(nvecs, veclen, ntypes, typeEnds, initStepSize, T)
    {
        if (rdist != nullptr) { this->setSRRDParams(SRRDDist<T>::params(*rdist)); }
    }
};
using GaussianAllMove = SRRDAllMoveOf<SRRDType::Gaussian>;
using StudentAllMove = SRRDAllMoveOf<SRRDType::Student>;
using CauchyAllMove = SRRDAllMoveOf<SRRDType::Cauchy>;
using ExponentialAllMove = SRRDAllMoveOf<SRRDType::Exponential>;
using GammaAllMove = SRRDAllMoveOf<SRRDType::Gamma>;
using WeibullAllMove = SRRDAllMoveOf<SRRDType::Weibull>;
using LognormalAllMove = SRRDAllMoveOf<SRRDType::Lognormal>;
using ChisqAllMove = SRRDAllMoveOf<SRRDType::Chisq>;
using FisherAllMove = SRRDAllMoveOf<SRRDType::Fisher>;
using GaussianVecMove = SRRDVecMoveOf<SRRDType::Gaussian>;
using StudentVecMove = SRRDVecMoveOf<SRRDType::Student>;
using CauchyVecMove = SRRDVecMoveOf<SRRDType::Cauchy>;
using ExponentialVecMove = SRRDVecMoveOf<SRRDType::Exponential>;
using GammaVecMove = SRRDVecMoveOf<SRRDType::Gamma>;
using WeibullVecMove = SRRDVecMoveOf<SRRDType::Weibull>;
using LognormalVecMove = SRRDVecMoveOf<SRRDType::Lognormal>;
using ChisqVecMove = SRRDVecMoveOf<SRRDType::Chisq>;
using FisherVecMove = SRRDVecMoveOf<SRRDType::Fisher>;

// MultiStepMove (include/mci/MultiStepMove.hpp:21-84, src/MultiStepMove.cpp:6-47): mini-Metropolis of nsteps sub-moves driven by
// the move's own sampling functions; default sub-move = uniform single-index move with step 0.05, default nsteps = ndim.
class MultiStepMove final: public TrialMoveInterface
{
    int _nsteps;
    std::unique_ptr<TrialMoveInterface> _trialMove;
    std::vector<std::unique_ptr<SamplingFunctionInterface>> _pdfs;
    TrialMoveInterface * _clone() const final
    {
        auto * ret = new MultiStepMove(_ndim, _nsteps);
        ret->setTrialMove(*_trialMove);
        for (auto & p : _pdfs) { ret->addSamplingFunction(*p); }
        return ret;
    }

public:
    MultiStepMove(int ndim, int nsteps): TrialMoveInterface(ndim), _nsteps(nsteps), _trialMove(new SRRDVecMove(ndim, 1, DEFAULT_MRT2STEP)) {}
    explicit MultiStepMove(int ndim): MultiStepMove(ndim, ndim) {}
    void setNSteps(int nsteps) { _nsteps = nsteps; }
    void setTrialMove(const TrialMoveInterface & tmove)
    {
        if (tmove.getNDim() != _ndim) {
            throw std::invalid_argument("[MultiStepMove::setTrialMove] Passed trial move's number of inputs is not equal to number of walkers.");
        }
        if (tmove.getMoveType() != MoveType::Vec) { throw std::domain_error("[MultiStepMove::setTrialMove] the device engine supports single-vector sub-moves only"); }
        _trialMove = tmove.clone();
    }
    void addSamplingFunction(const SamplingFunctionInterface & pdf)
    {
        if (pdf.getNDim() != _ndim) {
            throw std::invalid_argument("[MultiStepMove::addSamplingFunction] Passed sampling function's number of inputs is not equal to number of walkers.");
        }
        _pdfs.emplace_back(pdf.clone());
    }
    void clearSamplingFunctions() { _pdfs.clear(); }
    int getNSteps() const { return _nsteps; }
    TrialMoveInterface & getTrialMove() const { return *_trialMove; }
    int getNPDF() const { return static_cast<int>(_pdfs.size()); }
    SamplingFunctionInterface & getSamplingFunction(int i) const { return *_pdfs[static_cast<size_t>(i)]; }

    MoveType getMoveType() const final { return MoveType::MultiStep; }
    int getVecLen() const final { return _trialMove->getVecLen(); }
    int getNTypes() const final { return _trialMove->getNTypes(); }
    const int * getTypeEnds() const final { return _trialMove->getTypeEnds(); }
    int getNStepSizes() const final { return _trialMove->getNStepSizes(); }
    double getStepSize(int i) const final { return _trialMove->getStepSize(i); }
    void setStepSize(int i, double val) final { _trialMove->setStepSize(i, val); }
    double getChangeRate() const final { return std::min(1., _trialMove->getChangeRate()*_nsteps); }
    int getStepSizeIndex(int xidx) const final { return _trialMove->getStepSizeIndex(xidx); }
};
} // namespace mci
#endif
