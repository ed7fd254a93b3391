// mcig_comm.cuh — the collective of a sharded job INSIDE the library (included by mcig_engine.cu after its error helpers).
//
// Replaces the reference's MPI layer on the data path (paths relative to the reference repository):
//   MPIMCI::init / myrank / size / finalize      src/MPIMCI.cpp:27-35, 95-104   -> mcig_comm_init_env / _rank / _size / _finalize
//   MPI_Allreduce(SUM) of [sum avg | sum err^2]   src/MPIMCI.cpp:85-87           -> ncclAllReduce on the engine's stream, on device memory
//   MPI_Allreduce(SUM) of the acceptance rate     src/MCIntegrator.cpp:131-138   -> same (u64 acceptance counts), inside the device-resident
//   MPI_Allreduce(SUM) of the equilibration data  src/MCIntegrator.cpp:21-34        calibration / decorrelation loops: no host hop
//
// One process per GPU (the reference's one-rank-per-core model); the communicator is process-global like MPI_COMM_WORLD.
// libnccl.so.2 is dlopen'ed lazily, so libmcig.so still loads (and every single-GPU path still runs) on a machine without NCCL.
// Rendezvous: rank 0 creates the ncclUniqueId; either the caller transports it (mcig_comm_get_unique_id + mcig_comm_init_rank: bench.py
// broadcasts it through torch.distributed) or the library does over TCP from the launcher's environment (mcig_comm_init_env:
// RANK / WORLD_SIZE / LOCAL_RANK / MASTER_ADDR / MASTER_PORT as torchrun and tools/mcirun.sh set them).
#pragma once

#include <arpa/inet.h>
#include <netdb.h>
#include <netinet/in.h>
#include <netinet/tcp.h>
#include <sys/socket.h>
#include <sys/types.h>

namespace mcig_comm {

// the few NCCL declarations used (nccl.h is not required to build the library)
typedef struct ncclComm * ncclComm_t;
struct ncclUniqueId { char internal[128]; };
enum { kNcclSuccess = 0, kNcclSum = 0, kNcclUint64 = 5, kNcclFloat64 = 8 };

struct NcclApi {
    void * h = nullptr;
    int (*GetUniqueId)(ncclUniqueId *) = nullptr;
    int (*CommInitRank)(ncclComm_t *, int, ncclUniqueId, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void *, void *, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    const char * (*GetErrorString)(int) = nullptr;
    int (*GetVersion)(int *) = nullptr;
};

inline NcclApi & nccl()
{
    static NcclApi n;
    static std::once_flag once;
    std::call_once(once, [] {
        // an already loaded libnccl.so.2 (e.g. the one torch brought into a Python process) is found by its soname
        std::vector<std::string> names;
        if (const char * e = getenv("MCIG_NCCL_LIB")) { names.push_back(e); }
        names.push_back("libnccl.so.2");
        names.push_back("/usr/lib/x86_64-linux-gnu/libnccl.so.2");
        names.push_back("libnccl.so");
        for (auto & nm : names) {
            n.h = dlopen(nm.c_str(), RTLD_NOW | RTLD_GLOBAL);
            if (n.h != nullptr) { break; }
        }
        if (n.h == nullptr) { return; }
#define MCIG_NSYM(field, sym) n.field = reinterpret_cast<decltype(n.field)>(dlsym(n.h, sym))
        MCIG_NSYM(GetUniqueId, "ncclGetUniqueId");
        MCIG_NSYM(CommInitRank, "ncclCommInitRank");
        MCIG_NSYM(CommDestroy, "ncclCommDestroy");
        MCIG_NSYM(AllReduce, "ncclAllReduce");
        MCIG_NSYM(GetErrorString, "ncclGetErrorString");
        MCIG_NSYM(GetVersion, "ncclGetVersion");
#undef MCIG_NSYM
    });
    if (n.h == nullptr || n.AllReduce == nullptr || n.CommInitRank == nullptr) {
        fail(MCIG_ERR_RUNTIME, "[mcig] libnccl.so.2 could not be loaded: a job sharded over several GPUs needs NCCL (set MCIG_NCCL_LIB to its path)");
    }
    return n;
}

inline void nccl_check(int rc, const char * what)
{
    if (rc != kNcclSuccess) {
        NcclApi & n = nccl();
        fail(MCIG_ERR_RUNTIME, std::string("[mcig] NCCL error in ") + what + ": " + (n.GetErrorString != nullptr ? n.GetErrorString(rc) : "?"));
    }
}

struct World {
    ncclComm_t comm = nullptr;
    int rank = 0, size = 1, device = 0;
    bool active() const { return comm != nullptr && size > 1; }
};

inline World & world()
{
    static World w;
    return w;
}

inline void init_rank(const void * unique_id, int rank, int nranks, int device)
{
    World & w = world();
    if (w.comm != nullptr) { fail(MCIG_ERR_RUNTIME, "[MPIMCI::init] MPI already initialized!"); } // src/MPIMCI.cpp:30-32
    if (nranks < 1 || rank < 0 || rank >= nranks || unique_id == nullptr) { fail(MCIG_ERR_INVALID_ARGUMENT, "[mcig_comm_init_rank] bad rank / size / id"); }
    CUDA_CHECK(cudaSetDevice(device));
    ncclUniqueId id;
    memcpy(&id, unique_id, sizeof(id));
    ncclComm_t c = nullptr;
    nccl_check(nccl().CommInitRank(&c, nranks, id, rank), "ncclCommInitRank");
    w.comm = c;
    w.rank = rank;
    w.size = nranks;
    w.device = device;
}

inline void finalize()
{
    World & w = world();
    if (w.comm != nullptr) {
        cudaSetDevice(w.device);
        cudaDeviceSynchronize();
        nccl().CommDestroy(w.comm);
    }
    w = World();
}

// ---- TCP rendezvous of the unique id (only used by init_env): rank 0 serves 128 bytes to each of the other ranks
inline int env_int(const char * name, int dflt)
{
    const char * e = getenv(name);
    return (e != nullptr && e[0] != 0) ? atoi(e) : dflt;
}

inline void send_all(int fd, const char * buf, size_t n)
{
    size_t off = 0;
    while (off < n) {
        const ssize_t k = send(fd, buf + off, n - off, MSG_NOSIGNAL);
        if (k <= 0) { fail(MCIG_ERR_RUNTIME, "[mcig_comm_init_env] rendezvous: send failed"); }
        off += (size_t)k;
    }
}

inline void recv_all(int fd, char * buf, size_t n)
{
    size_t off = 0;
    while (off < n) {
        const ssize_t k = recv(fd, buf + off, n - off, 0);
        if (k <= 0) { fail(MCIG_ERR_RUNTIME, "[mcig_comm_init_env] rendezvous: connection closed before the NCCL id arrived"); }
        off += (size_t)k;
    }
}

inline void exchange_id(ncclUniqueId & id, int rank, int nranks)
{
    const char * addr_env = getenv("MASTER_ADDR");
    const std::string addr = (addr_env != nullptr && addr_env[0] != 0) ? addr_env : "127.0.0.1";
    // torchrun's own store listens on MASTER_PORT: the id travels on a neighbouring port (MCIG_COMM_PORT overrides)
    const int port = env_int("MCIG_COMM_PORT", env_int("MASTER_PORT", 29500) + 17);
    const uint32_t magic = 0x4d434947u; // "MCIG"
    if (rank == 0) {
        const int ls = socket(AF_INET, SOCK_STREAM, 0);
        if (ls < 0) { fail(MCIG_ERR_RUNTIME, "[mcig_comm_init_env] rendezvous: socket() failed"); }
        int one = 1;
        setsockopt(ls, SOL_SOCKET, SO_REUSEADDR, &one, sizeof(one));
        sockaddr_in sa;
        memset(&sa, 0, sizeof(sa));
        sa.sin_family = AF_INET;
        sa.sin_addr.s_addr = htonl(INADDR_ANY);
        sa.sin_port = htons((uint16_t)port);
        if (bind(ls, reinterpret_cast<sockaddr *>(&sa), sizeof(sa)) != 0 || listen(ls, nranks) != 0) {
            close(ls);
            fail(MCIG_ERR_RUNTIME, "[mcig_comm_init_env] rendezvous: cannot listen on port " + std::to_string(port) + " (set MCIG_COMM_PORT)");
        }
        for (int k = 1; k < nranks; ++k) {
            const int fd = accept(ls, nullptr, nullptr);
            if (fd < 0) { close(ls); fail(MCIG_ERR_RUNTIME, "[mcig_comm_init_env] rendezvous: accept() failed"); }
            uint32_t hello = 0;
            recv_all(fd, reinterpret_cast<char *>(&hello), sizeof(hello));
            if (hello != magic) { close(fd); --k; continue; } // a stray connection: ignore
            send_all(fd, id.internal, sizeof(id.internal));
            close(fd);
        }
        close(ls);
    }
    else {
        addrinfo hints, * res = nullptr;
        memset(&hints, 0, sizeof(hints));
        hints.ai_family = AF_INET;
        hints.ai_socktype = SOCK_STREAM;
        if (getaddrinfo(addr.c_str(), std::to_string(port).c_str(), &hints, &res) != 0 || res == nullptr) {
            fail(MCIG_ERR_RUNTIME, "[mcig_comm_init_env] rendezvous: cannot resolve MASTER_ADDR " + addr);
        }
        int fd = -1;
        const auto t0 = std::chrono::steady_clock::now();
        for (;;) { // rank 0 may not be listening yet
            fd = socket(AF_INET, SOCK_STREAM, 0);
            if (fd >= 0 && connect(fd, res->ai_addr, res->ai_addrlen) == 0) { break; }
            if (fd >= 0) { close(fd); }
            fd = -1;
            if (std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count() > 120.) { break; }
            usleep(20000);
        }
        freeaddrinfo(res);
        if (fd < 0) { fail(MCIG_ERR_RUNTIME, "[mcig_comm_init_env] rendezvous: rank 0 not reachable at " + addr + ":" + std::to_string(port)); }
        send_all(fd, reinterpret_cast<const char *>(&magic), sizeof(magic));
        recv_all(fd, id.internal, sizeof(id.internal));
        close(fd);
    }
}

// MPIMCI::init(): rank / size / device from the launcher's environment, id over TCP. A single process (no WORLD_SIZE) is a world of one
// and never touches NCCL.
inline int init_env()
{
    World & w = world();
    if (w.comm != nullptr || w.size > 1) { fail(MCIG_ERR_RUNTIME, "[MPIMCI::init] MPI already initialized!"); }
    const int nranks = env_int("WORLD_SIZE", env_int("OMPI_COMM_WORLD_SIZE", 1));
    const int rank = env_int("RANK", env_int("OMPI_COMM_WORLD_RANK", 0));
    if (nranks <= 1) {
        w.rank = 0;
        w.size = 1;
        return 0;
    }
    int ndev = 0;
    if (cudaGetDeviceCount(&ndev) != cudaSuccess || ndev == 0) {
        cudaGetLastError();
        fail(MCIG_ERR_CUDA, "[mcig] no CUDA device available: this library has no CPU fallback (sm_100a only)");
    }
    const int device = env_int("LOCAL_RANK", env_int("OMPI_COMM_WORLD_LOCAL_RANK", rank))%ndev;
    ncclUniqueId id;
    memset(&id, 0, sizeof(id));
    if (rank == 0) { nccl_check(nccl().GetUniqueId(&id), "ncclGetUniqueId"); }
    exchange_id(id, rank, nranks);
    init_rank(&id, rank, nranks, device);
    return rank;
}

inline void allreduce_sum_f64(double * dptr, size_t n, cudaStream_t stream)
{
    World & w = world();
    if (!w.active() || n == 0) { return; }
    nccl_check(nccl().AllReduce(dptr, dptr, n, kNcclFloat64, kNcclSum, w.comm, stream), "ncclAllReduce(f64)");
}

inline void allreduce_sum_u64(unsigned long long * dptr, size_t n, cudaStream_t stream)
{
    World & w = world();
    if (!w.active() || n == 0) { return; }
    nccl_check(nccl().AllReduce(dptr, dptr, n, kNcclUint64, kNcclSum, w.comm, stream), "ncclAllReduce(u64)");
}

} // namespace mcig_comm
