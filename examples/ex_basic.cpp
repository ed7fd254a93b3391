// ex_basic — the reference's examples/ex_basic/main.cpp workflow written against this repository's include/mci facade:
// integrate (4x - x^2) over [-1,3] (exact 20/3), first by plain sampling of the box, then with the sampling function
// g(x) = |x|/5. The only change a user of the reference makes is the include path of the fixture functions.
//   g++ -std=c++14 -Iinclude examples/ex_basic.cpp -Lmcintegratorplusplus_b200 -lmcig -Wl,-rpath,$PWD/mcintegratorplusplus_b200
#include <iostream>
#include <memory>

#include "mci/DeviceFunctions.hpp"
#include "mci/MCIntegrator.hpp"

int main(int argc, char ** argv)
{
    using namespace std;
    using namespace mci;
    const int64_t nwalkers = argc > 1 ? atoll(argv[1]) : 1024;

    const int ndim = 1;
    MCI mci(ndim);
    mci.setSeed(1337);
    mci.setNWalkers(nwalkers); // engine extension: independent chains, combined like MPI ranks
    mci.setIRange(-1., 3.);
    double initpos[ndim] = {-0.5};
    mci.setX(initpos);
    mci.setMRT2Step(0.25);
    mci.setTargetAcceptanceRate(0.7);

    Parabola obs;
    mci.addObservable(obs);
    const int Nmc = 100000;
    double average[1], error[1];
    mci.integrate(Nmc, average, error);
    cout << "no sampling function : " << average[0] << " +- " << error[0] << "  (exact 6.66667)" << endl;
    const bool ok1 = fabs(average[0] - 20./3.) < 5*error[0];

    mci.clearObservables();
    mci.addObservable(std::make_unique<NormalizedParabola>());
    std::unique_ptr<SamplingFunctionInterface> pdf = std::make_unique<NormalizedLine>();
    mci.addSamplingFunction(std::move(pdf));
    mci.integrate(Nmc, average, error);
    cout << "with g(x) = |x|/5     : " << average[0] << " +- " << error[0] << "  acceptance " << mci.getAcceptanceRate() << " step " << mci.getMRT2Step(0) << endl;
    const bool ok2 = fabs(average[0] - 20./3.) < 5*error[0];
    cout << (ok1 && ok2 ? "OK" : "MISMATCH") << endl;
    return (ok1 && ok2) ? 0 : 1;
}
