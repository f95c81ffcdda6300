#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/bench_sparse.py c5 512 front > gpurun_out/r2_c5_512_front19_pl8.log 2>&1; echo "pl8"; grep "lm_it\|solve (gram\|numeric\|substitutions" gpurun_out/r2_c5_512_front19_pl8.log | cut -c1-160
THB200_LIB=$PWD/scratch/libs/libthb200_pl4.so timeout 300 python scratch/bench_sparse.py c5 512 front > gpurun_out/r2_c5_512_front19_pl4.log 2>&1; echo "pl4"; grep "lm_it\|solve (gram\|numeric\|substitutions" gpurun_out/r2_c5_512_front19_pl4.log | cut -c1-160
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_c5_512_front19_launches.csv python scratch/prof_sparse.py c5 512 front > gpurun_out/r2_prof19.log 2>&1
python scratch/agg_launches.py gpurun_out/r2_c5_512_front19_launches.csv 90 > gpurun_out/r2_c5_512_front19_agg.txt 2>&1; head -18 gpurun_out/r2_c5_512_front19_agg.txt | cut -c1-175
timeout 600 python -m pytest tests/test_gpu_front.py -m gpu -x -q --timeout=500 -p no:cacheprovider --tb=short > gpurun_out/r2_front_tests19.log 2>&1; echo "front tests rc=$?"; tail -4 gpurun_out/r2_front_tests19.log
