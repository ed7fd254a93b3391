// Factories — mirror of include/mci/Factories.hpp: estimator selection and move factories (the accumulator factory's
// blocksize rule 0 -> Simple, 1 -> Full, >1 -> Block is applied inside the engine, mcig_add_obs).
#ifndef MCIG_MCI_FACTORIES_HPP
#define MCIG_MCI_FACTORIES_HPP

#include "mci/Estimators.hpp"
#include "mci/TrialMoveInterface.hpp"

#include <functional>
#include <memory>
#include <stdexcept>

namespace mci
{
enum class EstimatorType { Noop, Uncorrelated, Correlated, FCBlocker, MJBlocker }; // include/mci/Factories.hpp:52-59

inline EstimatorType selectEstimatorType(bool flag_correlated, bool flag_error = true)
{ // include/mci/Factories.hpp:61-71
    if (flag_correlated) {
        if (!flag_error) { throw std::invalid_argument("[selectEstimatorType] Error calculation is set off, but correlated error estimation is set on."); }
        return EstimatorType::Correlated;
    }
    return flag_error ? EstimatorType::Uncorrelated : EstimatorType::Noop;
}

inline std::function<void(int64_t, int, const double[], double[], double[])> createEstimator(EstimatorType estimType)
{ // include/mci/Factories.hpp:73-95
    switch (estimType) {
    case EstimatorType::Noop: return NoopEstimator;
    case EstimatorType::Uncorrelated: return UncorrelatedEstimator;
    case EstimatorType::Correlated: return CorrelatedEstimator;
    case EstimatorType::FCBlocker: return FCBlockerEstimator;
    case EstimatorType::MJBlocker: return MJBlockerEstimator;
    default: throw std::domain_error("[createEstimator] Unhandled estimator enumerator.");
    }
}

inline void checkTrialMoveSanity(int ndim, int ntypes = 1, const int typeEnds[] = nullptr)
{
    if (ndim < 1) { throw std::invalid_argument("[checkTrialMoveSanity] ndim must be at least 1."); }
    if (ntypes > 1 && typeEnds == nullptr) { throw std::invalid_argument("[checkTrialMoveSanity] ntypes>1 requires passed typeEnds array."); }
}

inline std::unique_ptr<TrialMoveInterface> createSRRDAllMove(SRRDType srrd, int ndim, int ntypes = 1, const int typeEnds[] = nullptr)
{
    ntypes = std::max(1, ntypes);
    checkTrialMoveSanity(ndim, ntypes, typeEnds);
    return std::unique_ptr<TrialMoveInterface>(new SRRDAllMove(ndim, ntypes, typeEnds, DEFAULT_MRT2STEP, srrd));
}

inline std::unique_ptr<TrialMoveInterface> createSRRDVecMove(SRRDType srrd, int nvecs, int veclen = 1, int ntypes = 1, const int typeEnds[] = nullptr)
{
    veclen = std::max(1, veclen);
    ntypes = std::max(1, ntypes);
    checkTrialMoveSanity(nvecs*veclen, ntypes, typeEnds);
    return std::unique_ptr<TrialMoveInterface>(new SRRDVecMove(nvecs, veclen, ntypes, typeEnds, DEFAULT_MRT2STEP, srrd));
}

inline std::unique_ptr<TrialMoveInterface> createMoveDefault(MoveType mtype, int ndim)
{ // include/mci/Factories.hpp:229-243
    switch (mtype) {
    case MoveType::All: return createSRRDAllMove(SRRDType::Uniform, ndim);
    case MoveType::Vec: return createSRRDVecMove(SRRDType::Uniform, ndim);
    case MoveType::MultiStep: return std::unique_ptr<TrialMoveInterface>(new MultiStepMove(ndim));
    default: throw std::domain_error("[createMoveDefault] Unhandled MoveType enumerator.");
    }
}
} // namespace mci
#endif
