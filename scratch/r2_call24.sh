#!/bin/bash
mkdir -p gpurun_out
nproc
timeout 1200 python bench.py > gpurun_out/r2_bench24.json 2> gpurun_out/r2_bench24.err; echo "bench rc=$?"; head -c 400 gpurun_out/r2_bench24.json; echo; tail -2 gpurun_out/r2_bench24.err
timeout 500 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_ref24.json 2> gpurun_out/r2_ref24.err; echo "ref rc=$?"; head -c 400 gpurun_out/r2_ref24.json; echo; tail -c 600 gpurun_out/r2_ref24.json
