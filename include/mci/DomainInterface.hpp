// Domains — mirror of include/mci/DomainInterface.hpp:12-56, UnboundDomain.hpp:12-37, OrthoPeriodicDomain.hpp/.cpp.
// The periodic wrap itself runs inside the walk kernel (device/mcig_device.cuh: OrthoPeriodicDomain); these host classes
// carry the bounds and answer the size/volume queries the reference's API exposes.
#ifndef MCIG_MCI_DOMAININTERFACE_HPP
#define MCIG_MCI_DOMAININTERFACE_HPP

#include "mci/Clonable.hpp"
#include "mci/DeviceFunctor.hpp"

#include <algorithm>
#include <limits>
#include <stdexcept>
#include <vector>

namespace mci
{
namespace domain_conv
{
static constexpr double infinity = std::numeric_limits<float>::max();
static constexpr double infinityX2 = infinity + infinity;
static constexpr double infiniteVol = 0.;
} // namespace domain_conv

struct DomainInterface: public Clonable<DomainInterface>
{
protected:
    explicit DomainInterface(int n_dim): ndim(n_dim)
    {
        if (ndim < 1) { throw std::invalid_argument("[DomainInterface] Number of dimensions must be at least 1."); }
    }

public:
    const int ndim;
    bool isFinite() const { return (this->getVolume() != domain_conv::infiniteVol); }
    void getCenter(double centerX[]) const
    {
        std::fill(centerX, centerX + ndim, 0.5);
        this->scaleToDomain(centerX);
    }
    // engine hooks. The two built-in domains are recognised by isPeriodic / getBounds. A user-defined domain (the reference's DomainInterface is
    // user-subclassable, include/mci/DomainInterface.hpp:26-55) implements the reference's five methods below for the host side and returns the
    // device twin of applyDomain / scaleToDomain from deviceFunctor() (plugin kind MCIG_PLUGIN_DOMAIN: wrap(i, x), scale(i, u01); include/mcig.h)
    virtual bool isPeriodic() const { return false; }
    virtual void getBounds(double lbounds[], double ubounds[]) const
    {
        std::fill(lbounds, lbounds + ndim, -domain_conv::infinity);
        std::fill(ubounds, ubounds + ndim, domain_conv::infinity);
    }
    virtual DeviceFunctor deviceFunctor() const { return DeviceFunctor(); }
    virtual void applyDomain(double x[]) const = 0;
    virtual void scaleToDomain(double normX[]) const = 0;
    virtual void getSizes(double dimSizes[]) const = 0;
    virtual double getVolume() const = 0;
};

struct UnboundDomain final: public DomainInterface
{
protected:
    DomainInterface * _clone() const final { return new UnboundDomain(ndim); }

public:
    explicit UnboundDomain(int n_dim): DomainInterface(n_dim) {}
    bool isPeriodic() const final { return false; }
    void getBounds(double lb[], double ub[]) const final
    {
        std::fill(lb, lb + ndim, -domain_conv::infinity);
        std::fill(ub, ub + ndim, domain_conv::infinity);
    }
    void applyDomain(double[]) const final {}
    void scaleToDomain(double normX[]) const final
    {
        for (int i = 0; i < ndim; ++i) { normX[i] = -domain_conv::infinity + normX[i]*domain_conv::infinityX2; }
    }
    void getSizes(double dimSizes[]) const final { std::fill(dimSizes, dimSizes + ndim, domain_conv::infinityX2); }
    double getVolume() const final { return domain_conv::infiniteVol; }
};

struct OrthoPeriodicDomain final: public DomainInterface
{
    std::vector<double> lbounds, ubounds;

protected:
    DomainInterface * _clone() const final { return new OrthoPeriodicDomain(ndim, lbounds.data(), ubounds.data()); }
    void _checkBounds() const
    {
        for (int i = 0; i < ndim; ++i) {
            if (ubounds[i] <= lbounds[i]) {
                throw std::invalid_argument("[OrthoPeriodicDomain::checkBounds] All upper bounds must be truly greater than their corresponding lower bounds.");
            }
        }
    }

public:
    explicit OrthoPeriodicDomain(int n_dim, double l_bound = -domain_conv::infinity, double u_bound = domain_conv::infinity):
            DomainInterface(n_dim), lbounds(n_dim, l_bound), ubounds(n_dim, u_bound) { _checkBounds(); }
    OrthoPeriodicDomain(int n_dim, const double l_bounds[], const double u_bounds[]):
            DomainInterface(n_dim), lbounds(l_bounds, l_bounds + n_dim), ubounds(u_bounds, u_bounds + n_dim) { _checkBounds(); }
    bool isPeriodic() const final { return true; }
    void getBounds(double lb[], double ub[]) const final
    {
        std::copy(lbounds.begin(), lbounds.end(), lb);
        std::copy(ubounds.begin(), ubounds.end(), ub);
    }
    void applyDomain(double x[]) const final
    {
        for (int i = 0; i < ndim; ++i) {
            const double len = ubounds[i] - lbounds[i];
            while (x[i] < lbounds[i]) { x[i] += len; }
            while (x[i] > ubounds[i]) { x[i] -= len; }
        }
    }
    void scaleToDomain(double normX[]) const final
    {
        for (int i = 0; i < ndim; ++i) { normX[i] = lbounds[i] + normX[i]*(ubounds[i] - lbounds[i]); }
    }
    void getSizes(double dimSizes[]) const final
    {
        for (int i = 0; i < ndim; ++i) { dimSizes[i] = ubounds[i] - lbounds[i]; }
    }
    double getVolume() const final
    {
        double vol = 1.;
        for (int i = 0; i < ndim; ++i) { vol *= (ubounds[i] - lbounds[i]); }
        return vol;
    }
};
} // namespace mci
#endif
