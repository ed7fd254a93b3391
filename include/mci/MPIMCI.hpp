// MPIMCI — the reference's optional MPI wrapper (include/mci/MPIMCI.hpp:14-29, src/MPIMCI.cpp) re-targeted: "ranks" are
// walkers of one MCI (and, in a multi-process job, the walkers of all processes). integrate() is therefore MCI::integrate
// itself: every walker runs the full Nmc and the results are combined as src/MPIMCI.cpp:85-92. A job sharded over several
// GPUs installs a cross-process sum with MCI::setAllreduce (see mcintegratorplusplus_b200/parallel.py for the
// torch.distributed/NCCL launcher); no MPI library is involved.
#ifndef MCIG_MCI_MPIMCI_HPP
#define MCIG_MCI_MPIMCI_HPP

#include "mci/MCIntegrator.hpp"

#include <fstream>
#include <string>
#include <vector>

namespace MPIMCI
{
inline int myrank() { return 0; }
inline int size(const mci::MCI & mci) { return static_cast<int>(mci.getNWalkers()); }
inline int init() { return 0; }
inline void finalize() {}

// walker w <- entry offset+w of a whitespace separated seed file (src/MPIMCI.cpp:38-65)
inline void setSeed(mci::MCI & mci, const std::string & filename, int offset = 0)
{
    std::ifstream seedfile(filename);
    if (!seedfile.good()) { throw std::runtime_error("Random seed file could not be found."); }
    for (int i = 0; i < offset; ++i) {
        if (seedfile.eof()) { throw std::runtime_error("Chosen seed offset is already larger than the number of seeds in seed file."); }
        uint_fast64_t skip;
        seedfile >> skip;
    }
    const int64_t n = mci.getNWalkers();
    std::vector<uint64_t> seeds(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) {
        if (seedfile.eof()) { throw std::runtime_error("Seed file doesn't provide enough seeds for the chosen number of ranks and offset."); }
        seedfile >> seeds[static_cast<size_t>(i)];
    }
    mci.setSeed(seeds[0]);
    mci.setWalkerSeeds(seeds.data(), n);
}

inline void integrate(mci::MCI & mci, int64_t Nmc, double average[], double error[], bool doFindMRT2Step = true, bool doDecorrelation = true)
{
    mci.integrate(Nmc, average, error, doFindMRT2Step, doDecorrelation);
}
} // namespace MPIMCI
#endif
