#!/bin/bash
mkdir -p gpurun_out
THB_BENCH_PROFILE_RANGE=1 timeout 170 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_bench_launches27.csv python bench.py --steps 1 --warmup 3 --no-c2 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu27.log 2>&1; echo "ncu bench rc=$?"
python scratch/agg_launches.py gpurun_out/r2_bench_launches27.csv 90 > gpurun_out/r2_bench_launches27_agg.txt 2>&1; head -24 gpurun_out/r2_bench_launches27_agg.txt | cut -c1-175
gzip -f gpurun_out/r2_bench_launches27.csv; ls -la gpurun_out/r2_bench_launches27.csv.gz
