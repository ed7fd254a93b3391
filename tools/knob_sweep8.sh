#!/bin/bash
run() {
  echo "== defs='$1'"
  MCIG_JIT_DEFINES="$1" python tools/profile_walk.py 100000 65536 0 1
  MCIG_JIT_DEFINES="$1" python tools/profile_walk.py 100000 303104 256 0
  MCIG_JIT_DEFINES="$1" python tools/profile_walk.py 100000 303104 0 1
}
run ""
run "MCIG_SPLIT_PROD=0"
run "MCIG_SPLIT_GROUP=1"
run "MCIG_SPLIT_GROUP=1;MCIG_WALK_UNROLL=1"
