#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/bench_sparse.py c5 512 front > gpurun_out/r2_c5_512_front10.log 2>&1; tail -6 gpurun_out/r2_c5_512_front10.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_c5_512_front10_launches.csv python scratch/prof_sparse.py c5 512 front > gpurun_out/r2_prof10.log 2>&1
python scratch/agg_launches.py gpurun_out/r2_c5_512_front10_launches.csv 90 > gpurun_out/r2_c5_512_front10_agg.txt 2>&1; head -16 gpurun_out/r2_c5_512_front10_agg.txt
timeout 2400 python -m pytest tests -m gpu -q --timeout=600 -p no:cacheprovider --tb=short -x > gpurun_out/r2_gpu_suite10.log 2>&1
echo "suite rc=$?" >> gpurun_out/r2_gpu_suite10.log; tail -15 gpurun_out/r2_gpu_suite10.log
