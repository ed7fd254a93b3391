// Reference-style SPMD program against the C++ facade (the call sequence of examples/ex_mpi/main.cpp:22-110): MPIMCI::init -> setSeed from a
// seed file -> MPIMCI::integrate with automatic step calibration and decorrelation -> finalize. Launched once per GPU by tools/mcirun.sh
// (or torchrun --no-python); rank 0 prints one line of results. tests/test_multi_gpu.py compares the N-process job with the same job run
// by one process over all walkers: sharding must not change the calibrated step size, and the estimates agree to rounding.
#include <cstdio>
#include <cstdlib>
#include <fstream>
#include <iostream>
#include <memory>

#include "mci/MCIntegrator.hpp"
#include "mci/MPIMCI.hpp"
#include "mci/DeviceFunctions.hpp"

int main(int argc, char ** argv)
{
    using namespace mci;
    const int64_t total_walkers = (argc > 1) ? atoll(argv[1]) : 4096;
    const int64_t nmc = (argc > 2) ? atoll(argv[2]) : 20000;
    const char * seedfile = (argc > 3) ? argv[3] : nullptr;
    const int myrank = MPIMCI::init();
    const int nranks = MPIMCI::size();
    if (total_walkers%nranks != 0) {
        std::cerr << "total walkers must be a multiple of the number of ranks" << std::endl;
        return 2;
    }
    MCI mci(3);
    mci.setRngMode(RngMode::Philox32);
    mci.setNWalkers(total_walkers/nranks);
    if (seedfile != nullptr) { MPIMCI::setSeed(mci, seedfile, 3); }
    else { mci.setSeed(5649871); }
    const double x0[3] = {1.0, -0.5, 0.25};
    mci.setX(x0);
    mci.setMRT2Step(0.2); // far from the target acceptance: findMRT2Step has to work
    mci.addSamplingFunction(ThreeDimGaussianPDF());
    mci.addObservable(XND(3), 0, 1);
    mci.addObservable(XSquared(), 1, 5);
    mci.addObservable(XYZSquared(), 5, 2);
    double average[7], error[7];
    MPIMCI::integrate(mci, nmc, average, error, true, true);
    if (myrank == 0) {
        printf("MPIMCI ranks %d step %.17g acc %.17g", nranks, mci.getMRT2Step(0), mci.getAcceptanceRate());
        for (int i = 0; i < 7; ++i) { printf(" %.17g %.17g", average[i], error[i]); }
        printf("\n");
    }
    // a second call continues the chains (positions and streams persist across integrate calls)
    MPIMCI::integrate(mci, nmc, average, error, false, false);
    if (myrank == 0) {
        printf("MPIMCI second");
        for (int i = 0; i < 7; ++i) { printf(" %.17g %.17g", average[i], error[i]); }
        printf("\n");
    }
    MPIMCI::finalize();
    return 0;
}
