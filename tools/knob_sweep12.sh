#!/bin/bash
# marginal accept path entered warp-uniformly (MCIG_ACCEPT_VOTE; the code was removed after this measurement, see accept_log in mcig_device.cuh)
run() {
  echo "== defs='$1'"
  MCIG_JIT_DEFINES="$1" python tools/profile_walk.py 100000 65536 0 1
  MCIG_JIT_DEFINES="$1" python tools/profile_walk.py 100000 303104 0 1
  MCIG_JIT_DEFINES="$1" python tools/profile_walk.py 100000 303104 0 0
}
run ""
run "MCIG_ACCEPT_VOTE=1"
run "MCIG_ACCEPT_VOTE=1;MCIG_WALK_UNROLL_DYN=2"
run ""
