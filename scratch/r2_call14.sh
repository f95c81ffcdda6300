#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/bench_sparse.py c5 512 front > gpurun_out/r2_c5_512_front14.log 2>&1; tail -6 gpurun_out/r2_c5_512_front14.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_c5_512_front14_launches.csv python scratch/prof_sparse.py c5 512 front > gpurun_out/r2_prof14.log 2>&1
python scratch/agg_launches.py gpurun_out/r2_c5_512_front14_launches.csv 90 > gpurun_out/r2_c5_512_front14_agg.txt 2>&1; head -12 gpurun_out/r2_c5_512_front14_agg.txt
timeout 600 python -m pytest tests/test_gpu_front.py -m gpu -q --timeout=500 -p no:cacheprovider --tb=short -x 2>&1 | tail -3
