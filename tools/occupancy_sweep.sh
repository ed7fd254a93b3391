#!/bin/bash
# per-warp step latency vs warps/scheduler: W = 148 SMs x 128 threads x k  (k warps per SMSP)
for k in 1 2 3 4 6 8 16; do timeout 60 python tools/profile_walk.py 20000 $((148*128*k)) 128 0; done
