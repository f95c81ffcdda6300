#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_front.py tests/test_gpu_gram_dense.py tests/test_gpu_linearize.py tests/test_gpu_sparse_solver.py tests/test_gpu_ba.py tests/test_gpu_extlib_kat.py -m gpu -q --timeout=500 -p no:cacheprovider --tb=short > gpurun_out/r2_tests8.log 2>&1
echo "rc=$?" >> gpurun_out/r2_tests8.log; tail -15 gpurun_out/r2_tests8.log
timeout 300 python scratch/bench_sparse.py c5 512 front > gpurun_out/r2_c5_512_front8.log 2>&1; tail -6 gpurun_out/r2_c5_512_front8.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_c5_512_front8_launches.csv python scratch/prof_sparse.py c5 512 front > gpurun_out/r2_prof8.log 2>&1
python scratch/agg_launches.py gpurun_out/r2_c5_512_front8_launches.csv 90 > gpurun_out/r2_c5_512_front8_agg.txt 2>&1; head -16 gpurun_out/r2_c5_512_front8_agg.txt
timeout 600 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:front_small_kernel --launch-skip 38 --launch-count 3 -o gpurun_out/r2_front_small_full python scratch/prof_sparse.py c5 512 front > gpurun_out/r2_prof8b.log 2>&1
timeout 600 ncu --set full --clock-control none --profile-from-start off -k regex:gram_block --launch-count 2 -o gpurun_out/r2_gram_block_full python scratch/prof_sparse.py c5 512 front > gpurun_out/r2_prof8c.log 2>&1
ls -la gpurun_out/*.ncu-rep
