#!/bin/bash
# all-moves beyond ndim 64: full unrolling / draws up front instead of the 4-fold unrolled streaming loop
for defs in "" "MCIG_UNROLL_MAX=128" "MCIG_STREAM_NDIM=128;MCIG_UNROLL_MAX=128" "MCIG_STREAM_NDIM=256;MCIG_UNROLL_MAX=256"; do
  echo "== defs='$defs'"
  MCIG_JIT_DEFINES="$defs" python -c "
import sys; sys.path.insert(0, 'tools'); import bench_suite as b
b.c3_ndim('all', ndims=(96, 128), nmc=2000)
if '256' in '$defs' or not '$defs': b.c3_ndim('all', ndims=(192, 256), nmc=1000)
" | cut -c1-130
done
