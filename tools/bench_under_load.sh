#!/bin/bash
# bench_under_load.sh OUT -- the headline bench line with every host core busy (2 spinning processes per core): the device-timed value must not depend on how
# quickly the host thread wakes up (one synchronisation per integrate call). Hogs are killed by PID.
out=$1
n=$(( $(nproc) * 2 ))
pids=()
for i in $(seq $n); do python -c "while True: pass" & pids+=($!); done
sleep 1
python bench.py --no-secondary --steps 20 > "$out" 2>/dev/null
for p in "${pids[@]}"; do kill $p 2>/dev/null; done
wait 2>/dev/null
