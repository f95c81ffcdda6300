#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_zz_first_run.py -m gpu -q --timeout=300 -p no:cacheprovider -k "so2_rotation or tactile" --tb=long > gpurun_out/r2_fail4.log 2>&1
echo "rc=$?" >> gpurun_out/r2_fail4.log
tail -30 gpurun_out/r2_fail4.log
timeout 300 python scratch/bench_sparse.py c5 512 lane_tiled_root > gpurun_out/r2_c5_512_ltr.log 2>&1; tail -8 gpurun_out/r2_c5_512_ltr.log
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_c5_512_ltr_launches.csv python scratch/prof_sparse.py c5 512 lane_tiled_root > gpurun_out/r2_prof.log 2>&1
python scratch/agg_launches.py gpurun_out/r2_c5_512_ltr_launches.csv 90 > gpurun_out/r2_c5_512_ltr_agg.txt 2>&1; head -30 gpurun_out/r2_c5_512_ltr_agg.txt
timeout 500 python scratch/bench_sparse.py c5 4096 lane_tiled_root > gpurun_out/r2_c5_4096_ltr.log 2>&1; tail -12 gpurun_out/r2_c5_4096_ltr.log
