// mci::DependentObservableInterface — observables that read the sampling functions' proto values and / or earlier observables
// (reference: include/mci/DependentObservableInterface.hpp:24-53).
//
// In the reference such an observable scans the containers in registerDeps() and keeps host pointers. On the device the
// dependencies are handed to the functor instead: a class that also derives from this interface returns a DeviceFunctor
// whose observableFunction takes a third argument,
//     template <class XV, class DEP> __device__ void observableFunction(const XV & x, double * out, const DEP & dep) const;
//     dep.proto(i)   i-th proto value of the sampling functions at the current position (flat over the pdfs, in the order added):
//                    what SamplingFunctionInterface::observationCallback(x, protovalues) is given in the reference
//     dep.obs(k, j)  j-th value of observable k as evaluated in this step (AccumulatorInterface::getObsValue)
// The reference's rules stay: (2) only observables at a lower index may be read, (3) their nskip must divide this one's
// (isObsDepValid below, same arithmetic). MCI sets DeviceFunctor::dependent for every observable deriving from this class.
#ifndef MCIG_MCI_DEPENDENTOBSERVABLEINTERFACE_HPP
#define MCIG_MCI_DEPENDENTOBSERVABLEINTERFACE_HPP

namespace mci
{
class DependentObservableInterface
{
protected:
    const bool _flag_pdfdep; // reads dep.proto(..)
    explicit DependentObservableInterface(bool dependsOnPDF): _flag_pdfdep(dependsOnPDF) {}

public:
    virtual ~DependentObservableInterface() = default;
    bool dependsOnPDF() const { return _flag_pdfdep; }
    static bool isObsDepValid(int selfIdx, int selfNskip, int depIdx, int depNskip)
    {
        const bool isOrdered = (depIdx < selfIdx);
        const bool isSynced = (selfNskip >= depNskip) ? (selfNskip%depNskip == 0) : false;
        return isOrdered && isSynced;
    }
};
} // namespace mci
#endif
