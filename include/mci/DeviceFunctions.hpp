// DeviceFunctions — the reference's fixture sampling functions / observables (test/common/TestMCIFunctions.hpp:123-473,
// examples/common/ExampleFunctions.hpp:11-82) as ready-made classes with the reference's names and constructor shapes.
// Their device functors are built into the engine (mcintegratorplusplus_b200/csrc/device/mcig_functors.cuh).
#ifndef MCIG_MCI_DEVICEFUNCTIONS_HPP
#define MCIG_MCI_DEVICEFUNCTIONS_HPP

#include "mci/ObservableFunctionInterface.hpp"
#include "mci/SamplingFunctionInterface.hpp"

#define MCIG_FIXED_PDF(NAME, NDIM, NPROTO)                                                        \
    class NAME final: public mci::SamplingFunctionInterface                                      \
    {                                                                                             \
        mci::SamplingFunctionInterface * _clone() const final { return new NAME(); }             \
                                                                                                  \
    public:                                                                                       \
        NAME(): mci::SamplingFunctionInterface(NDIM, NPROTO) {}                                   \
        mci::DeviceFunctor deviceFunctor() const final { return mci::DeviceFunctor(#NAME); }      \
    };
#define MCIG_ND_PDF(NAME)                                                                         \
    class NAME final: public mci::SamplingFunctionInterface                                      \
    {                                                                                             \
        mci::SamplingFunctionInterface * _clone() const final { return new NAME(_ndim); }        \
                                                                                                  \
    public:                                                                                       \
        explicit NAME(int ndim): mci::SamplingFunctionInterface(ndim, ndim) {}                    \
        mci::DeviceFunctor deviceFunctor() const final { return mci::DeviceFunctor(#NAME); }      \
    };
#define MCIG_FIXED_OBS(NAME, NDIM, NOBS)                                                          \
    class NAME final: public mci::ObservableFunctionInterface                                    \
    {                                                                                             \
        mci::ObservableFunctionInterface * _clone() const final { return new NAME(); }           \
                                                                                                  \
    public:                                                                                       \
        NAME(): mci::ObservableFunctionInterface(NDIM, NOBS, false) {}                            \
        mci::DeviceFunctor deviceFunctor() const final { return mci::DeviceFunctor(#NAME); }      \
    };
#define MCIG_ND_OBS(NAME, NOBS_EXPR, UPD)                                                         \
    class NAME final: public mci::ObservableFunctionInterface                                    \
    {                                                                                             \
        mci::ObservableFunctionInterface * _clone() const final { return new NAME(_ndim); }      \
                                                                                                  \
    public:                                                                                       \
        explicit NAME(int ndim): mci::ObservableFunctionInterface(ndim, NOBS_EXPR, UPD) {}        \
        mci::DeviceFunctor deviceFunctor() const final { return mci::DeviceFunctor(#NAME); }      \
    };

MCIG_FIXED_PDF(ThreeDimGaussianPDF, 3, 1) // exp(-(x^2+y^2+z^2))
MCIG_ND_PDF(Gauss)                        // exp(-sum x_i^2), selective update
MCIG_FIXED_PDF(Exp1DPDF, 1, 1)            // exp(-|x|)
MCIG_ND_PDF(ExpNDPDF)                     // exp(-sum |x_i|), selective update
MCIG_FIXED_PDF(NormalizedLine, 1, 1)      // |x|/5 on [-1,3]

MCIG_FIXED_OBS(XSquared, 3, 1)
MCIG_FIXED_OBS(GaussXSquared, 3, 1)
MCIG_FIXED_OBS(XYZSquared, 3, 3)
MCIG_FIXED_OBS(X1D, 1, 1)
MCIG_ND_OBS(XND, ndim, false)
MCIG_ND_OBS(UpdateableXND, ndim, true)
MCIG_ND_OBS(Constval, 1, false)
MCIG_ND_OBS(Polynom, 1, false)
MCIG_ND_OBS(X2Sum, 1, false)
MCIG_ND_OBS(X2, ndim, true)
MCIG_FIXED_OBS(Parabola, 1, 1)
MCIG_FIXED_OBS(NormalizedParabola, 1, 1)

#undef MCIG_FIXED_PDF
#undef MCIG_ND_PDF
#undef MCIG_FIXED_OBS
#undef MCIG_ND_OBS
#endif
