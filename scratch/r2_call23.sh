#!/bin/bash
mkdir -p gpurun_out
timeout 300 python __graft_entry__.py --smoke > gpurun_out/r2_smoke23.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2_smoke23.log | cut -c1-200
timeout 900 python -m pytest tests/test_gpu_front.py tests/test_gpu_extlib_kat.py tests/test_gpu_pgo_benchmark.py tests/test_gpu_c4_tactile.py tests/test_gpu_gram_dense.py -m gpu -x -q --timeout=600 -p no:cacheprovider --tb=short > gpurun_out/r2_tests23.log 2>&1; echo "tests rc=$?"; tail -3 gpurun_out/r2_tests23.log
timeout 900 python bench.py > gpurun_out/r2_bench23.json 2> gpurun_out/r2_bench23.err; echo "bench rc=$?"; head -c 400 gpurun_out/r2_bench23.json; echo; tail -2 gpurun_out/r2_bench23.err
timeout 400 python bench.py --impl reference --steps 1 --warmup 0 > gpurun_out/r2_ref23.json 2> gpurun_out/r2_ref23.err; echo "ref rc=$?"; head -c 500 gpurun_out/r2_ref23.json
