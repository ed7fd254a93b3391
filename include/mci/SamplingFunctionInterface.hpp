// SamplingFunctionInterface — mirror of include/mci/SamplingFunctionInterface.hpp:36-105 / ProtoFunctionInterface.hpp:24-65.
// Users derive from it exactly as in the reference (constructor (ndim, nproto), protected _clone()); instead of overriding
// protoFunction / samplingFunction / acceptanceFunction / updatedAcceptance on the host they return the device functor
// implementing the same four methods (see DeviceFunctor.hpp).
#ifndef MCIG_MCI_SAMPLINGFUNCTIONINTERFACE_HPP
#define MCIG_MCI_SAMPLINGFUNCTIONINTERFACE_HPP

#include "mci/Clonable.hpp"
#include "mci/DeviceFunctor.hpp"

namespace mci
{
class ProtoFunctionInterface
{
protected:
    const int _ndim;
    int _nproto;
    ProtoFunctionInterface(int ndim, int nproto): _ndim(ndim), _nproto(nproto)
    {
        if (ndim < 1) { throw std::invalid_argument("[ProtoFunctionInterface] Number of dimensions must be at least 1."); }
    }

public:
    virtual ~ProtoFunctionInterface() = default;
    int getNDim() const { return _ndim; }
    int getNProto() const { return _nproto; }
};

class SamplingFunctionInterface: public ProtoFunctionInterface, public Clonable<SamplingFunctionInterface>
{
protected:
    SamplingFunctionInterface(int ndim, int nproto): ProtoFunctionInterface(ndim, nproto) {}

public:
    // the __device__ functor with protoFunction / samplingFunction / acceptanceFunction (/ updatedAcceptance)
    virtual DeviceFunctor deviceFunctor() const = 0;
};
} // namespace mci
#endif
