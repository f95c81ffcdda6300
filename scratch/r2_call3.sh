#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_front.py -m gpu -q --timeout=400 -p no:cacheprovider --tb=short > gpurun_out/r2_front_tests3.log 2>&1
echo "rc=$?" >> gpurun_out/r2_front_tests3.log; tail -15 gpurun_out/r2_front_tests3.log
timeout 300 python scratch/bench_sparse.py c5 512 front > gpurun_out/r2_c5_512_front3.log 2>&1; tail -9 gpurun_out/r2_c5_512_front3.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_c5_512_front3_launches.csv python scratch/prof_sparse.py c5 512 front > gpurun_out/r2_prof3.log 2>&1
python scratch/agg_launches.py gpurun_out/r2_c5_512_front3_launches.csv 90 > gpurun_out/r2_c5_512_front3_agg.txt 2>&1; head -20 gpurun_out/r2_c5_512_front3_agg.txt
grep "front_small_kernel\|chol_col_kernel\|assemble" gpurun_out/r2_c5_512_front3_launches.csv | awk -F'","' '{print $5, $(NF)}' | sed 's/"//g' | awk '{print $1, $NF}' | head -80 > gpurun_out/r2_per_launch3.txt
