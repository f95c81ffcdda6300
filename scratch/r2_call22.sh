#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python scratch/bench_sparse.py c5 512 front > gpurun_out/r2_c5_512_front22_$tag.log 2>&1; echo "$tag rc=$?"; grep "lm_it\|numeric\|substitutions\|Error\|error" gpurun_out/r2_c5_512_front22_$tag.log | cut -c1-130; }
run la1 THB_X=0
run la0 THB200_LIB=$PWD/scratch/libs/libthb200_la0.so
timeout 1200 python -m pytest tests/ -x -q -m gpu -p no:cacheprovider --tb=short > gpurun_out/r2_gpu_suite22.log 2>&1; echo "suite rc=$?"; tail -3 gpurun_out/r2_gpu_suite22.log
timeout 900 python bench.py > gpurun_out/r2_bench22.json 2> gpurun_out/r2_bench22.err; echo "bench rc=$?"; head -c 600 gpurun_out/r2_bench22.json; echo; tail -3 gpurun_out/r2_bench22.err
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_c5_512_front22_launches.csv python scratch/prof_sparse.py c5 512 front > gpurun_out/r2_prof22.log 2>&1
python scratch/agg_launches.py gpurun_out/r2_c5_512_front22_launches.csv 90 > gpurun_out/r2_c5_512_front22_agg.txt 2>&1; head -12 gpurun_out/r2_c5_512_front22_agg.txt | cut -c1-175
timeout 300 ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name-base demangled -k "regex:front_small_kernel<\(int\)1024>" -s 7 -c 1 -f -o gpurun_out/r2_fs1024_la python scratch/prof_sparse.py c5 512 front > gpurun_out/r2_ncu22.log 2>&1; echo "ncu rc=$?"
