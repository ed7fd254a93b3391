// ObservableFunctionInterface — mirror of include/mci/ObservableFunctionInterface.hpp:30-63 (constructor (ndim, nobs,
// isUpdateable), protected _clone()); observableFunction lives in the device functor. An observable is a pure function of
// the walker position, so the reference's updatedObservable optimisation is value-neutral and not needed on the device.
#ifndef MCIG_MCI_OBSERVABLEFUNCTIONINTERFACE_HPP
#define MCIG_MCI_OBSERVABLEFUNCTIONINTERFACE_HPP

#include "mci/Clonable.hpp"
#include "mci/DeviceFunctor.hpp"

namespace mci
{
class ObservableFunctionInterface: public Clonable<ObservableFunctionInterface>
{
protected:
    const int _ndim;
    const int _nobs;
    const bool _flag_updateable;
    ObservableFunctionInterface(int ndim, int nobs, bool isUpdateable): _ndim(ndim), _nobs(nobs), _flag_updateable(isUpdateable) {}

public:
    int getNObs() const { return _nobs; }
    int getNDim() const { return _ndim; }
    bool isUpdateable() const { return _flag_updateable; }
    virtual DeviceFunctor deviceFunctor() const = 0;
};
} // namespace mci
#endif
