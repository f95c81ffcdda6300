#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_front.py tests/test_gpu_extlib_kat.py tests/test_gpu_sparse_solver.py -m gpu -q --timeout=500 -p no:cacheprovider --tb=short > gpurun_out/r2_tests9.log 2>&1
echo "rc=$?" >> gpurun_out/r2_tests9.log; tail -12 gpurun_out/r2_tests9.log
timeout 300 python scratch/bench_sparse.py c5 512 front > gpurun_out/r2_c5_512_front9.log 2>&1; tail -6 gpurun_out/r2_c5_512_front9.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_c5_512_front9_launches.csv python scratch/prof_sparse.py c5 512 front > gpurun_out/r2_prof9.log 2>&1
python scratch/agg_launches.py gpurun_out/r2_c5_512_front9_launches.csv 90 > gpurun_out/r2_c5_512_front9_agg.txt 2>&1; head -16 gpurun_out/r2_c5_512_front9_agg.txt
timeout 1500 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench9.json 2> gpurun_out/r2_bench9.err; echo "bench rc=$?"; head -c 1500 gpurun_out/r2_bench9.json; echo; tail -5 gpurun_out/r2_bench9.err
