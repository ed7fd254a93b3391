#!/bin/bash
# second-session knobs of the register-resident walk loop: observable cache, counter asm, unroll of the dynamically scheduled kernel
run() {
  echo "== defs='$1' obs_cache=$2"
  MCIG_JIT_DEFINES="$1" MCIG_OBS_CACHE=$2 python tools/profile_walk.py 100000 65536 0 1
  MCIG_JIT_DEFINES="$1" MCIG_OBS_CACHE=$2 python tools/profile_walk.py 100000 65536 512 0
  MCIG_JIT_DEFINES="$1" MCIG_OBS_CACHE=$2 python tools/profile_walk.py 100000 303104 256 0
}
run "" 1
run "" 0
run "MCIG_NACC_ASM=0" 1
run "MCIG_WALK_UNROLL_DYN=2" 1
run "MCIG_WALK_UNROLL_DYN=2;MCIG_NACC_ASM=0" 1
run "MCIG_WALK_UNROLL=1" 1
run "MCIG_WALK_UNROLL=4;MCIG_WALK_UNROLL_DYN=4" 1
run "MCIG_ACCEPT_OUTLINE=1" 1
