#!/bin/bash
# mcirun.sh N prog [args...] — start N copies of a reference-style SPMD program (MPIMCI::init / integrate / finalize), one per GPU of this
# node: the mpirun of the NCCL-based MPIMCI (include/mci/MPIMCI.hpp). Sets RANK / LOCAL_RANK / WORLD_SIZE / MASTER_ADDR / MASTER_PORT, which
# mcig_comm_init_env reads; the programs exchange the NCCL id over TCP on MASTER_PORT + 17. Exit code: the first non-zero one.
set -u
N=$1; shift
export WORLD_SIZE=$N MASTER_ADDR=${MASTER_ADDR:-127.0.0.1} MASTER_PORT=${MASTER_PORT:-29511}
pids=()
for ((r = 0; r < N; ++r)); do
    RANK=$r LOCAL_RANK=$r "$@" &
    pids+=($!)
done
rc=0
for p in "${pids[@]}"; do
    wait "$p" || rc=$?
done
exit $rc
