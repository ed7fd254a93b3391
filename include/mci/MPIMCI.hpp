// MPIMCI — the reference's MPI wrapper (include/mci/MPIMCI.hpp:14-29, src/MPIMCI.cpp:11-104) for one process per GPU.
//
// Same free functions, same call sequence (examples/ex_mpi/main.cpp): init() -> setSeed(mci, file) -> integrate(mci, Nmc, avg, err) ->
// finalize(). The "MPI" underneath is the library's own NCCL communicator (include/mcig.h: mcig_comm_*): init() reads rank / size /
// device from the launcher's environment (RANK, WORLD_SIZE, LOCAL_RANK as set by torchrun or tools/mcirun.sh; OMPI_COMM_WORLD_* under
// mpirun) and exchanges the NCCL id over TCP; a program started alone is a world of one process and never loads NCCL.
//
// Mapping of the reference's model: there one rank = one MCI = one chain. Here every process runs the MCI's W walkers on its GPU, so a
// job of R processes samples R x W chains; rank r owns the global walker ids [r W, (r+1) W). integrate() attaches the integrator to
// the communicator: the rank-averaged acceptance rate of findMRT2Step (src/MCIntegrator.cpp:131-138), the equilibration estimates of
// initialDecorrelation (:21-34) and the final [sum avg | sum err^2] (src/MPIMCI.cpp:85-92) are ncclAllReduce calls on the engine's
// stream, inside the device-resident control loops; the result is the reference's combination with every walker counted as a rank:
// average = sum_w avg_w / (R W), error = sqrt(sum_w err_w^2) / (R W).
#ifndef MCIG_MCI_MPIMCI_HPP
#define MCIG_MCI_MPIMCI_HPP

#include "mci/MCIntegrator.hpp"

#include <cstdint>
#include <fstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace MPIMCI
{
namespace detail
{
inline bool & initialized()
{
    static bool f = false;
    return f;
}
inline bool & finalized()
{
    static bool f = false;
    return f;
}
inline void requireActive()
{ // src/MPIMCI.cpp:71-75, 98-102
    if (!initialized()) { throw std::runtime_error("MPI not initialized!"); }
    if (finalized()) { throw std::runtime_error("MPI already finalized!"); }
}
} // namespace detail

// return my rank (not named rank(): it would collide with "using namespace std", as the reference notes)
inline int myrank() { return mcig_comm_rank(); }

// return size (number of processes = GPUs of the job)
inline int size() { return mcig_comm_size(); }

// init the communicator and return the rank of this process (src/MPIMCI.cpp:27-35)
inline int init()
{
    if (detail::initialized()) { throw std::runtime_error("MPI already initialized!"); }
    const int rank = mcig_comm_init_env();
    if (rank < 0) { throw std::runtime_error(mcig_last_error()); }
    detail::initialized() = true;
    return rank;
}

// Different random seeds per chain from a whitespace separated file (src/MPIMCI.cpp:38-65): the reference hands entry offset + rank to
// each rank; here global walker g = rank W + w takes entry offset + g, so that R processes x W walkers use the seeds R W single-walker
// ranks would. (Philox modes key the streams by the first seed and the global walker id; replay mode seeds one std::mt19937_64 per walker.)
inline void setSeed(mci::MCI & mci, const std::string & filename, int offset = 0)
{
    std::ifstream seedfile(filename);
    if (!seedfile.good()) { throw std::runtime_error("Random seed file could not be found."); }
    const int64_t n = mci.getNWalkers();
    const int64_t first = static_cast<int64_t>(offset) + static_cast<int64_t>(myrank())*n;
    for (int64_t i = 0; i < first; ++i) {
        if (seedfile.eof()) { throw std::runtime_error("Chosen seed offset is already larger than the number of seeds in seed file."); }
        uint_fast64_t skip;
        seedfile >> skip;
    }
    std::vector<uint64_t> seeds(static_cast<size_t>(n));
    for (int64_t i = 0; i < n; ++i) {
        if (seedfile.eof()) { throw std::runtime_error("Seed file doesn't provide enough seeds for the chosen number of ranks and offset."); }
        seedfile >> seeds[static_cast<size_t>(i)];
    }
    // Philox key: the job-wide first entry (every rank reads the same one), so that all shards draw from ONE family of streams
    std::ifstream again(filename);
    uint_fast64_t key = 0;
    for (int i = 0; i <= offset; ++i) { again >> key; }
    mci.setSeed(key);
    mci.setWalkerSeeds(seeds.data(), n);
}

// integrate in parallel and accumulate results (src/MPIMCI.cpp:68-93): every process samples its shard, the combination runs on the GPUs
inline void integrate(mci::MCI & mci, int64_t Nmc, double average[], double error[], bool doFindMRT2Step = true, bool doDecorrelation = true)
{
    detail::requireActive();
    const int64_t n = mci.getNWalkers();
    const int r = myrank(), R = size();
    mci.setNWalkers(n, static_cast<int64_t>(r)*n, static_cast<int64_t>(R)*n); // this rank's shard of the R x W chains
    mci.attachComm(R > 1);
    mci.integrate(Nmc, average, error, doFindMRT2Step, doDecorrelation);
}

// finalize the communicator (src/MPIMCI.cpp:95-104)
inline void finalize()
{
    detail::requireActive();
    if (mcig_comm_finalize() != 0) { throw std::runtime_error(mcig_last_error()); }
    detail::finalized() = true;
}
} // namespace MPIMCI
#endif
