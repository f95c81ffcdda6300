#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_front.py tests/test_gpu_gram_dense.py tests/test_gpu_extlib_kat.py tests/test_gpu_c4_tactile.py tests/test_gpu_pgo_benchmark.py -m gpu -q --timeout=500 -p no:cacheprovider --tb=short -k "not sparse_lane" > gpurun_out/r2_tests7.log 2>&1
echo "rc=$?" >> gpurun_out/r2_tests7.log; tail -25 gpurun_out/r2_tests7.log
timeout 300 python scratch/bench_sparse.py c5 512 front > gpurun_out/r2_c5_512_front7.log 2>&1; tail -6 gpurun_out/r2_c5_512_front7.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_c5_512_front7_launches.csv python scratch/prof_sparse.py c5 512 front > gpurun_out/r2_prof7.log 2>&1
python scratch/agg_launches.py gpurun_out/r2_c5_512_front7_launches.csv 90 > gpurun_out/r2_c5_512_front7_agg.txt 2>&1; head -14 gpurun_out/r2_c5_512_front7_agg.txt
for cfg in "0.2 200000 128" "0.3 400000 160"; do
  set -- $cfg
  THB_FRONT_TAU=$1 THB_FRONT_MERGE_FLOPS=$2 THB_FRONT_MERGE_MAX_R=$3 timeout 200 python scratch/bench_sparse.py c5 512 front > gpurun_out/r2_c5_tune7_$1_$2_$3.log 2>&1
  echo "tune $cfg:"; grep "solve (gram\|numeric\|substitutions" gpurun_out/r2_c5_tune7_$1_$2_$3.log | cut -c1-200
done
