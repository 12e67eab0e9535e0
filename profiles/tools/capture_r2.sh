#!/bin/bash
# Round-2 evidence, run on the GPU box through gpurun (one GPU):
#   bash profiles/tools/capture_r2.sh
# Outputs go to gpurun_out/r2x/; the summaries that should be judged are copied into profiles/r2/.
OUT=gpurun_out/r2x
mkdir -p $OUT
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/smoke.log 2>&1
python bench.py --steps 20 --warmup 3 > $OUT/bench.json 2> $OUT/bench.err
# launch list of the bench command (cold-cache, serialised times: shares, not absolutes)
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file $OUT/launches_bench.csv \
    python bench.py --steps 3 --warmup 3 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
# full captures of the dominant kernels
ncu --set full --import-source on --clock-control none -k regex:fused_loop --launch-skip 2 -c 1 -f -o $OUT/fused_full \
    python profiles/tools/time_fused.py --reps 1 > $OUT/ncu_fused.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:tc_encoder --launch-skip 3 -c 1 -f -o $OUT/tc_full \
    python profiles/tools/time_tc_lift.py 4000000 20000 > $OUT/ncu_tc.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:encoder_units --launch-skip 3 -c 1 -f -o $OUT/enc_units_full \
    python profiles/tools/time_tc_lift.py 4000000 20000 > $OUT/ncu_enc.log 2>&1
ncu --set full --clock-control none -k regex:gram_dmma --launch-skip 3 -c 1 -f -o $OUT/gram_full \
    python profiles/tools/time_tc_lift.py 4000000 20000 > $OUT/ncu_gram.log 2>&1
ncu --set full --import-source on --clock-control none -k regex:loop_qp_plant --launch-skip 125 -c 1 -f -o $OUT/tank_qp_full \
    python profiles/tools/tank_profile_workload.py > $OUT/ncu_tank.log 2>&1
# probes
(cd profiles/tools && nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_mix_probe fp64_mix_probe.cu) && ./profiles/tools/fp64_mix_probe > $OUT/fp64_mix_probe.txt 2>&1
python profiles/tools/time_fused_chunks.py 4096 10 > $OUT/fused_chunk_profile.txt 2>&1
python profiles/tools/time_tank_modes.py 65536 300 > $OUT/tank_modes.json 2> $OUT/tank_modes.err
python profiles/tools/time_tc_lift.py > $OUT/time_tc.json 2> $OUT/time_tc.err
tail -2 $OUT/smoke.log
ncu --set full --import-source on --clock-control none -k regex:loop_qp_plant --launch-skip 45 -c 1 -f -o $OUT/rbf50_qp_full \
    python profiles/tools/rbf50_profile_workload.py > $OUT/ncu_rbf50.log 2>&1
for m in 0 3 2 1; do python profiles/tools/rbf50_profile_workload.py 125000 100 $m | tail -1; done > $OUT/rbf50_modes.txt 2>&1
