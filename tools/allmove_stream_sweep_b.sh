#!/bin/bash
# all-moves in state memory after the footprint change: draws up front (Draws<D>) vs block by block (StreamDraws), unroll caps
for defs in "" "MCIG_STREAM_NDIM=16" "MCIG_STREAM_NDIM=16;MCIG_UNROLL_MAX=16" "MCIG_UNROLL_MAX=16" "MCIG_UNROLL_MAX=32"; do
  echo "== defs='$defs'"
  MCIG_JIT_DEFINES="$defs" python -c "
import sys; sys.path.insert(0, 'tools'); import bench_suite as b
b.c3_ndim('all', ndims=(32, 48, 64), nmc=5000)
b.c3_ndim('all', ndims=(128,), nmc=1000)
" | cut -c1-130
done
