#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_front.py -m gpu -q --timeout=400 -p no:cacheprovider --tb=short > gpurun_out/r2_front_tests5.log 2>&1
echo "rc=$?" >> gpurun_out/r2_front_tests5.log; tail -6 gpurun_out/r2_front_tests5.log
timeout 300 python scratch/bench_sparse.py c5 512 front > gpurun_out/r2_c5_512_front5.log 2>&1; tail -7 gpurun_out/r2_c5_512_front5.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_c5_512_front5_launches.csv python scratch/prof_sparse.py c5 512 front > gpurun_out/r2_prof5.log 2>&1
python scratch/agg_launches.py gpurun_out/r2_c5_512_front5_launches.csv 90 > gpurun_out/r2_c5_512_front5_agg.txt 2>&1; head -14 gpurun_out/r2_c5_512_front5_agg.txt
THB_BENCH_C5_BATCH=512 timeout 900 python bench.py --steps 2 --warmup 3 > gpurun_out/r2_bench5.json 2> gpurun_out/r2_bench5.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/r2_bench5.json; tail -5 gpurun_out/r2_bench5.err
