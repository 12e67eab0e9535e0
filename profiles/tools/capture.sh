#!/bin/bash
# ncu evidence for the bench command (run on the GPU box through gpurun):
#   1. launch list of `bench.py --steps 3 --warmup 3` (gpu__time_duration per launch)
#   2. one --set full capture of the fused closed-loop kernel (details page + raw metrics)
# Outputs go to gpurun_out/; copy what should be judged into profiles/rNN/.
set -x
OUT=gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:fused_loop --launch-skip 2 -c 1 -f -o $OUT/fused_full \
    python profiles/tools/time_fused.py --reps 1 > $OUT/ncu_full.log 2>&1
ncu -i $OUT/fused_full.ncu-rep --page details > $OUT/ncu_full_fused_details.txt 2>/dev/null
ncu -i $OUT/fused_full.ncu-rep --page raw --csv > $OUT/ncu_full_fused_raw.csv 2>/dev/null
