#!/bin/bash
# per-warp step latency vs warps/scheduler: W = 148 SMs x 128 threads x k  (k warps per SMSP)
nvidia-smi --query-gpu=clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 200 > gpurun_out/clocks_sweep.csv &
SMI=$!
for k in 1 2 3 4 6 8 16; do python tools/profile_walk.py 20000 $((148*128*k)) 128; done
python tools/profile_walk.py 100000 65536 128
python tools/profile_walk.py 100000 65536 512
kill $SMI
sort gpurun_out/clocks_sweep.csv | uniq -c | sort -rn | head -8
