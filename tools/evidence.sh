#!/bin/bash
# One gpurun call that refreshes the evidence the judge reads (run from the repo root ON the GPU box):
#   gpurun --timeout 900 -- 'bash tools/evidence.sh r02a'
# writes gpurun_out/<tag>_*: GPU test log, smoke, bench lines (ours + reference arm), ncu launch list of the bench command, ncu --set full of the
# headline walk kernel (raw csv + per-instruction stall samples of its loop). Copy what should be judged into profiles/.
tag=${1:-rXX}
out=gpurun_out
mkdir -p $out
(time timeout 600 python -m pytest tests -m gpu -x -q) > $out/${tag}_pytest_gpu.log 2>&1
python -c "import __graft_entry__ as g; g.smoke()" >> $out/${tag}_pytest_gpu.log 2>&1
tail -4 $out/${tag}_pytest_gpu.log
timeout 300 python bench.py > $out/${tag}_bench_1gpu.json 2> $out/${tag}_bench_1gpu.err
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > $out/${tag}_bench_reference_arm.json 2>> $out/${tag}_bench_1gpu.err
timeout 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -c 400 --csv \
    --log-file $out/${tag}_bench_launches_ncu.csv python bench.py --steps 2 --warmup 3 --no-secondary > /dev/null 2>&1
timeout 300 ncu --set full --section SourceCounters --clock-control none --import-source on -k regex:mcig_walk_dyn --launch-skip 2 --launch-count 1 \
    -f -o $out/${tag}_walk python tools/profile_walk.py 3000 65536 0 1 > $out/${tag}_ncu_walk.log 2>&1
ncu -i $out/${tag}_walk.ncu-rep --page raw --csv > $out/${tag}_walk_ncu_raw.csv 2>/dev/null
ncu -i $out/${tag}_walk.ncu-rep --page source --csv --print-source sass > $out/${tag}_walk_ncu_source.csv 2>/dev/null
tail -c 300 $out/${tag}_bench_1gpu.json
