#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python scratch/bench_sparse.py c5 512 front > gpurun_out/r2_c5_512_front21_$tag.log 2>&1; echo "$tag rc=$?"; grep "lm_it\|numeric\|substitutions\|Error\|error" gpurun_out/r2_c5_512_front21_$tag.log | cut -c1-130; }
run pdl1 THB_X=0
run pdl0 THB_FRONT_PDL=0
run pdl1_graph THB_X=0 BENCH_GRAPH=1
timeout 600 python -m pytest tests/test_gpu_front.py tests/test_gpu_extlib_kat.py tests/test_gpu_pgo_benchmark.py -m gpu -x -q --timeout=500 -p no:cacheprovider --tb=short > gpurun_out/r2_front_tests21.log 2>&1; echo "front tests rc=$?"; tail -4 gpurun_out/r2_front_tests21.log
