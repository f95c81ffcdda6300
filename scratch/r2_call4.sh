#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_front.py tests/test_gpu_pgo_benchmark.py -m gpu -q --timeout=400 -p no:cacheprovider --tb=short -s > gpurun_out/r2_front_tests4.log 2>&1
echo "rc=$?" >> gpurun_out/r2_front_tests4.log; grep "pgo_benchmark\[" gpurun_out/r2_front_tests4.log; tail -8 gpurun_out/r2_front_tests4.log
timeout 300 python scratch/bench_sparse.py c5 512 front > gpurun_out/r2_c5_512_front4.log 2>&1; tail -9 gpurun_out/r2_c5_512_front4.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_c5_512_front4_launches.csv python scratch/prof_sparse.py c5 512 front > gpurun_out/r2_prof4.log 2>&1
python scratch/agg_launches.py gpurun_out/r2_c5_512_front4_launches.csv 90 > gpurun_out/r2_c5_512_front4_agg.txt 2>&1; head -16 gpurun_out/r2_c5_512_front4_agg.txt
