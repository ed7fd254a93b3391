#!/bin/bash
# seed-specialised kernels: Philox round keys as LOP3 immediates (MCIG_RK_IMM) with one / two / four steps per loop trip
run() {
  echo "== imm=$1 defs='$2'"
  MCIG_TOOL_RK_IMM=$1 MCIG_JIT_DEFINES="$2" python tools/profile_walk.py 100000 65536 0 1
  MCIG_TOOL_RK_IMM=$1 MCIG_JIT_DEFINES="$2" python tools/profile_walk.py 100000 303104 0 1
  MCIG_TOOL_RK_IMM=$1 MCIG_JIT_DEFINES="$2" python tools/profile_walk.py 100000 303104 0 0
}
run 0 ""
run 1 ""
run 1 "MCIG_WALK_UNROLL_DYN=2"
run 1 "MCIG_WALK_UNROLL=1"
run 1 "MCIG_WALK_UNROLL=4;MCIG_WALK_UNROLL_DYN=4"
