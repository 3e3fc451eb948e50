#!/bin/bash
# Round evidence in one GPU call: bench line (with the CPU baseline), ncu launch list of the same command, and one
# `ncu --set full` capture of the two blend kernels (+ the small kernels).  Outputs land in gpurun_out/prof/; copy the
# summaries into profiles/ afterwards (tools/summarise_profiles.py).
set -u
OUT=gpurun_out/prof
mkdir -p $OUT
python bench.py > $OUT/bench.json 2> $OUT/bench.err
python bench.py --impl reference --steps 20 --warmup 1 > $OUT/bench_reference.json 2>> $OUT/bench.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $OUT/launches.csv \
    python bench.py --steps 2 --warmup 1 --no-cpu-baseline > $OUT/bench_under_ncu.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:blend_ -s 4 -c 2 -o $OUT/blend \
    python tools/prof_fwd.py bwd > $OUT/ncu_blend.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"preprocess|plan_kernel|scatter|sort_" -s 12 -c 6 -o $OUT/small \
    python tools/prof_fwd.py bwd > $OUT/ncu_small.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:"knn_|prep_cov3d" -s 22 -c 11 -o $OUT/prep \
    python tools/prof_prep.py > $OUT/ncu_prep.log 2>&1
ncu -i $OUT/prep.ncu-rep --page details > $OUT/prep_details.txt 2>/dev/null
ncu -i $OUT/blend.ncu-rep --page details > $OUT/blend_details.txt 2>/dev/null
ncu -i $OUT/small.ncu-rep --page details > $OUT/small_details.txt 2>/dev/null
ncu -i $OUT/blend.ncu-rep --page raw --csv > $OUT/blend_raw.csv 2>/dev/null
tail -2 $OUT/bench.err
