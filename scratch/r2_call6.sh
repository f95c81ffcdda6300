#!/bin/bash
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_front.py tests/test_gpu_gram_dense.py tests/test_gpu_extlib_kat.py tests/test_gpu_c4_tactile.py -m gpu -q --timeout=500 -p no:cacheprovider --tb=short > gpurun_out/r2_tests6.log 2>&1
echo "rc=$?" >> gpurun_out/r2_tests6.log; tail -40 gpurun_out/r2_tests6.log
timeout 300 python scratch/bench_sparse.py c5 512 front > gpurun_out/r2_c5_512_front6.log 2>&1; tail -6 gpurun_out/r2_c5_512_front6.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_c5_512_front6_launches.csv python scratch/prof_sparse.py c5 512 front > gpurun_out/r2_prof6.log 2>&1
python scratch/agg_launches.py gpurun_out/r2_c5_512_front6_launches.csv 90 > gpurun_out/r2_c5_512_front6_agg.txt 2>&1; head -14 gpurun_out/r2_c5_512_front6_agg.txt
for cfg in "0.2 200000 128" "0.3 400000 160" "0.12 40000 96 64"; do
  set -- $cfg
  THB_FRONT_TAU=$1 THB_FRONT_MERGE_FLOPS=$2 THB_FRONT_MERGE_MAX_R=$3 THB_FRONT_STRIPE_BUDGET_KB=${4:-110} timeout 200 python scratch/bench_sparse.py c5 512 front > gpurun_out/r2_c5_tune_$1_$2_$3_${4:-110}.log 2>&1
  echo "tune $cfg:"; grep "symbolic\|solve (gram\|numeric\|substitutions" gpurun_out/r2_c5_tune_$1_$2_$3_${4:-110}.log | cut -c1-260
done
