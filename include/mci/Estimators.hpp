// Estimators — the reference's free estimator functions (include/mci/Estimators.hpp:9-45) with identical signatures,
// executed by the device kernels K3/K4/K5 (host/mcig_kernels.cuh) through mcig_estimate. Data layout as in the reference:
// x[n*ndim], sample-major.
#ifndef MCIG_MCI_ESTIMATORS_HPP
#define MCIG_MCI_ESTIMATORS_HPP

#include "mci/DeviceFunctor.hpp"

#include <cstdint>

namespace mci
{
inline void NoopEstimator(int64_t n, int ndim, const double x[], double average[], double error[]) { detail::check(mcig_estimate(MCIG_EST_NOOP, n, ndim, x, average, error)); }
inline void UncorrelatedEstimator(int64_t n, int ndim, const double x[], double average[], double error[]) { detail::check(mcig_estimate(MCIG_EST_UNCORRELATED, n, ndim, x, average, error)); }
inline void CorrelatedEstimator(int64_t n, int ndim, const double x[], double average[], double error[]) { detail::check(mcig_estimate(MCIG_EST_CORRELATED, n, ndim, x, average, error)); }
inline void FCBlockerEstimator(int64_t n, int ndim, const double x[], double average[], double error[]) { detail::check(mcig_estimate(MCIG_EST_FCBLOCKER, n, ndim, x, average, error)); }
inline void MJBlockerEstimator(int64_t n, int ndim, const double x[], double average[], double error[]) { detail::check(mcig_estimate(MCIG_EST_MJBLOCKER, n, ndim, x, average, error)); }
inline void MultiDimUncorrelatedEstimator(int64_t n, int ndim, const double x[], double average[], double error[]) { UncorrelatedEstimator(n, ndim, x, average, error); }
inline void MultiDimFCBlockerEstimator(int64_t n, int ndim, const double x[], double average[], double error[]) { FCBlockerEstimator(n, ndim, x, average, error); }
inline void OneDimUncorrelatedEstimator(int64_t n, const double x[], double & average, double & error) { UncorrelatedEstimator(n, 1, x, &average, &error); }
inline void MultiDimBlockEstimator(int64_t n, int ndim, const double x[], int64_t nblocks, double average[], double error[]) { detail::check(mcig_estimate_blocks(n, ndim, x, nblocks, average, error)); }
inline void OneDimBlockEstimator(int64_t n, const double x[], int64_t nblocks, double & average, double & error) { detail::check(mcig_estimate_blocks(n, 1, x, nblocks, &average, &error)); }
inline void OneDimFCBlockerEstimator(int64_t n, const double x[], double & average, double & error) { FCBlockerEstimator(n, 1, x, &average, &error); }
} // namespace mci
#endif
