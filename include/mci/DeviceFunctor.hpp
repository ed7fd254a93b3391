// DeviceFunctor — what a "plugin" is in the B200 engine.
//
// In the reference a sampling function / observable is a host class with virtual methods that MCI calls once per Metropolis
// step (include/mci/SamplingFunctionInterface.hpp:36-105, ObservableFunctionInterface.hpp:30-63). A virtual host call per
// step cannot exist inside a GPU kernel, so here the same class hierarchy carries a *device functor*: the name of a
// registered __device__ functor type, (optionally) its CUDA source, and its run-time parameters. MCI hands these to the
// C-ABI (mcig_register_plugin / mcig_add_pdf / mcig_add_obs), which JIT-specialises the walk kernel for them.
// The functor contract is documented in mcintegratorplusplus_b200/csrc/device/mcig_functors.cuh.
#ifndef MCIG_MCI_DEVICEFUNCTOR_HPP
#define MCIG_MCI_DEVICEFUNCTOR_HPP

#include "mcig.h"

#include <stdexcept>
#include <string>
#include <vector>

namespace mci
{
namespace detail
{
inline void check(int rc)
{ // error codes back to the exceptions the reference throws
    if (rc == MCIG_OK) { return; }
    const std::string msg = mcig_last_error();
    switch (rc) {
    case MCIG_ERR_INVALID_ARGUMENT: throw std::invalid_argument(msg);
    case MCIG_ERR_DOMAIN: throw std::domain_error(msg);
    default: throw std::runtime_error(msg);
    }
}
} // namespace detail

struct DeviceFunctor
{
    std::string name;      // registry key (the reference's fixtures are pre-registered under their class names)
    std::string typeExpr;  // C++ type of the functor, "{ndim}" is substituted; empty = already registered under `name`
    std::string source;    // CUDA C++ defining typeExpr (empty for built-ins)
    std::vector<double> params; // run-time parameters (functor member `par`)
    bool hasUpdate = false;     // functor overrides updatedAcceptance
    bool elementwise = false;   // proto value k depends on x[k] only
    bool logAcceptance = false; // functor provides logAcceptance(protoold, protonew)
    bool dependent = false;     // observable functor takes observableFunction(x, out, dep) (mci::DependentObservableInterface)

    DeviceFunctor() = default;
    explicit DeviceFunctor(std::string registeredName, std::vector<double> par = {}): name(std::move(registeredName)), params(std::move(par)) {}
    DeviceFunctor(std::string n, std::string type, std::string src, std::vector<double> par = {}, bool upd = false, bool elem = false, bool logacc = false):
            name(std::move(n)), typeExpr(std::move(type)), source(std::move(src)), params(std::move(par)), hasUpdate(upd), elementwise(elem), logAcceptance(logacc) {}

    // returns the plugin id, registering the functor first if it brings its own source
    int resolve(int kind, int ndim, int nvalues) const
    {
        if (!typeExpr.empty()) {
            const int flags = (hasUpdate ? MCIG_PLUGIN_HAS_UPDATE : 0) | (elementwise ? MCIG_PLUGIN_ELEMENTWISE : 0) | (logAcceptance ? MCIG_PLUGIN_LOG_ACCEPTANCE : 0) |
                              (dependent ? MCIG_PLUGIN_DEPENDENT : 0);
            const int id = mcig_register_plugin(kind, name.c_str(), typeExpr.c_str(), source.c_str(), ndim, nvalues, static_cast<int>(params.size()), flags);
            if (id < 0) { detail::check(-id); }
            return id;
        }
        const int id = mcig_lookup_plugin(kind, name.c_str());
        if (id < 0) { throw std::invalid_argument("[mci::DeviceFunctor] no device functor registered under the name " + name); }
        return id;
    }
};
} // namespace mci
#endif
