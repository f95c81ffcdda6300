#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_front.py -m gpu -q --timeout=400 -p no:cacheprovider --tb=short > gpurun_out/r2_front_tests.log 2>&1
echo "rc=$?" >> gpurun_out/r2_front_tests.log; tail -40 gpurun_out/r2_front_tests.log
timeout 600 python -m pytest tests/test_gpu_zz_first_run.py tests/test_gpu_dense_solver.py tests/test_gpu_lm.py -m gpu -q --timeout=300 -p no:cacheprovider -k "so2_rotation or tactile or dense_solver or test_gpu_lm" --tb=short > gpurun_out/r2_fix4.log 2>&1
echo "rc=$?" >> gpurun_out/r2_fix4.log; tail -40 gpurun_out/r2_fix4.log
timeout 300 python scratch/bench_sparse.py c5 512 front > gpurun_out/r2_c5_512_front.log 2>&1; tail -9 gpurun_out/r2_c5_512_front.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_c5_512_front_launches.csv python scratch/prof_sparse.py c5 512 front > gpurun_out/r2_prof2.log 2>&1
python scratch/agg_launches.py gpurun_out/r2_c5_512_front_launches.csv 90 > gpurun_out/r2_c5_512_front_agg.txt 2>&1; head -24 gpurun_out/r2_c5_512_front_agg.txt
timeout 200 python scratch/bench_sparse.py c5 128 front > gpurun_out/r2_c5_128_front.log 2>&1; tail -9 gpurun_out/r2_c5_128_front.log
