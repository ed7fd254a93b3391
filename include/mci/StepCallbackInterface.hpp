// mci::StepCallbackInterface — MCI::setCallback on the device (reference: include/mci/MCIntegrator.hpp:186-190).
//
// The reference calls a std::function<void(const MCI &)> after every move (src/MCIntegrator.cpp:267, :346, :374): host code
// inside the step loop. The device-resident walk calls a __device__ functor at the same three places instead,
//     template <class XO, class XN> __device__ void operator()(const XO & xold, const XN & xnew, bool accepted,
//                                                              long long walker, long long step, double * buffer) const;
// (step = -1 for the call from initializeSampling), with `buffer` = bufferDoubles() doubles of device memory that MCI zeroes
// at the start of every integrate call and hands back through MCI::getCallbackBuffer(). Walkers run concurrently: index the
// buffer by walker or use atomicAdd.
#ifndef MCIG_MCI_STEPCALLBACKINTERFACE_HPP
#define MCIG_MCI_STEPCALLBACKINTERFACE_HPP

#include "mci/Clonable.hpp"
#include "mci/DeviceFunctor.hpp"

#include <cstdint>

namespace mci
{
class StepCallbackInterface: public Clonable<StepCallbackInterface>
{
public:
    virtual DeviceFunctor deviceFunctor() const = 0;
    virtual int64_t bufferDoubles(int64_t nwalkers) const = 0;
};
} // namespace mci
#endif
