#!/bin/bash
for defs in "" "MCIG_SPLIT_GROUP=0"; do
  echo "== $defs"
  MCIG_JIT_DEFINES="$defs" python tools/profile_walk.py 100000 65536 0 1
  MCIG_JIT_DEFINES="$defs" python tools/profile_walk.py 100000 65536 512 0
  MCIG_JIT_DEFINES="$defs" python tools/profile_walk.py 100000 303104 256 0
done
