"""Multi-GPU layer: the reference's MPIMCI (src/MPIMCI.cpp) re-targeted at one process per GPU + torch.distributed.

The path shards by independent units: every walker is an independent chain, exactly the reference's MPI-rank model
(src/MPIMCI.cpp:83 runs the full Nmc on every rank). So walkers are partitioned over ranks with NO data-path collective;
the only exchange is the all-reduce of [sum_w avg_w | sum_w err_w^2] at the end of integrate (and, when the automatic
calibration / decorrelation are on, the tiny rate / estimate sums of src/MCIntegrator.cpp:135 and :21-34) — NCCL over
NVLink on GPUs, gloo in the CPU tests. Philox streams are keyed by the GLOBAL walker id, so per-walker results do not
depend on the number of GPUs.
"""
import numpy as np


def shard(total_walkers, rank, world):
    """Contiguous block partition of global walker ids: rank r owns [offset, offset + n)."""
    base, rem = divmod(int(total_walkers), int(world))
    n = base + (1 if rank < rem else 0)
    offset = rank*base + min(rank, rem)
    return n, offset


_staging = {}  # device index -> (pinned host tensor, device tensor): the messages are <= a few hundred bytes, allocating per call costs more than sending


def allreduce_sum(buf, group=None):
    """In-place sum of a numpy float64 array over all ranks (no-op without an initialised process group)."""
    import torch
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return buf
    if dist.get_backend(group) == "nccl":
        n = int(buf.size)
        dev = torch.cuda.current_device()
        st = _staging.get(dev)
        if st is None or st[0].numel() < n:
            cap = max(256, n)
            st = (torch.empty(cap, dtype=torch.float64).pin_memory(), torch.empty(cap, dtype=torch.float64, device="cuda"))
            _staging[dev] = st
        host, devt = st
        host[:n] = torch.from_numpy(np.ascontiguousarray(buf).reshape(-1))
        devt[:n].copy_(host[:n], non_blocking=True)
        dist.all_reduce(devt[:n], group=group)
        host[:n].copy_(devt[:n], non_blocking=True)
        torch.cuda.current_stream().synchronize()
        buf.reshape(-1)[:] = host[:n].numpy()
    else:
        t = torch.from_numpy(np.ascontiguousarray(buf).copy())
        dist.all_reduce(t, group=group)
        buf[:] = t.numpy()
    return buf


def combine(sums, nobsdim, total_walkers):
    """MPIMCI::integrate's normalisation (src/MPIMCI.cpp:89-92) of the all-reduced [sum avg | sum err^2]."""
    sums = np.asarray(sums, dtype=np.float64)
    avg = sums[:nobsdim]/float(total_walkers)
    err = np.sqrt(sums[nobsdim:2*nobsdim])/float(total_walkers)
    return avg, err


def init_comm(device=None):
    """Create the library's own NCCL communicator (include/mcig.h: mcig_comm_*) for the ranks of the initialised torch.distributed job:
    rank 0 draws the ncclUniqueId, torch.distributed only transports its 128 bytes. Afterwards MCI.attachComm() makes every all-reduce of
    the sampling path an ncclAllReduce on the engine's stream, inside the device-resident control loops. Returns (rank, world)."""
    import ctypes as C
    import torch
    import torch.distributed as dist
    from . import _capi
    lib = _capi.lib()
    if not dist.is_initialized() or dist.get_world_size() == 1:
        return 0, 1
    if lib.mcig_comm_size() > 1:
        return lib.mcig_comm_rank(), lib.mcig_comm_size()
    rank, world = dist.get_rank(), dist.get_world_size()
    if device is None:
        device = torch.cuda.current_device()
    buf = C.create_string_buffer(128)
    if rank == 0:
        _capi.check(lib.mcig_comm_get_unique_id(buf))
    t = torch.frombuffer(bytearray(buf.raw), dtype=torch.uint8).clone()
    if dist.get_backend() == "nccl":
        t = t.cuda()
    dist.broadcast(t, 0)
    raw = bytes(t.cpu().numpy().tobytes())
    _capi.check(lib.mcig_comm_init_rank(C.create_string_buffer(raw, 128), rank, world, int(device)))
    return rank, world


def finalize_comm():
    from . import _capi
    _capi.check(_capi.lib().mcig_comm_finalize())


def install(mci, total_walkers, group=None):
    """Give `mci` its shard of `total_walkers` and the cross-rank sum. Returns (n_local, offset). On GPUs (NCCL backend, default group)
    the sum is the library's own ncclAllReduce (init_comm + attachComm); otherwise (gloo in the CPU tests, sub-groups) a host callback."""
    import torch.distributed as dist
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    n, off = shard(total_walkers, rank, world)
    mci.setNWalkers(n, global_offset=off, total=total_walkers)
    if world > 1:
        if group is None and dist.get_backend() == "nccl":
            init_comm()
            mci.attachComm()
        else:
            mci.setAllreduce(lambda buf: allreduce_sum(buf, group))
    return n, off


class MPIMCI:
    """Namespace mirror of include/mci/MPIMCI.hpp:14-29 for torch.distributed jobs."""

    @staticmethod
    def init(backend=None):
        import os
        import torch
        import torch.distributed as dist
        if not dist.is_initialized() and int(os.environ.get("WORLD_SIZE", "1")) > 1:
            backend = backend or ("nccl" if torch.cuda.is_available() else "gloo")
            if backend == "nccl":
                torch.cuda.set_device(int(os.environ.get("LOCAL_RANK", 0)))
            dist.init_process_group(backend)
        return MPIMCI.myrank()

    @staticmethod
    def myrank():
        import torch.distributed as dist
        return dist.get_rank() if dist.is_initialized() else 0

    @staticmethod
    def size():
        import torch.distributed as dist
        return dist.get_world_size() if dist.is_initialized() else 1

    @staticmethod
    def integrate(mci, Nmc, doFindMRT2Step=True, doDecorrelation=True):
        return mci.integrate(Nmc, doFindMRT2Step, doDecorrelation)

    @staticmethod
    def finalize():
        import torch.distributed as dist
        if dist.is_initialized():
            dist.destroy_process_group()
