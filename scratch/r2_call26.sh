#!/bin/bash
mkdir -p gpurun_out
timeout 420 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_bench_launches26.csv python bench.py --steps 1 --warmup 1 --no-c2 --no-cpu-baseline > gpurun_out/r2_bench_under_ncu26.log 2>&1; echo "ncu bench rc=$?"
python scratch/agg_launches.py gpurun_out/r2_bench_launches26.csv 90 > gpurun_out/r2_bench_launches26_agg.txt 2>&1; head -14 gpurun_out/r2_bench_launches26_agg.txt | cut -c1-175
