#!/bin/bash
mkdir -p gpurun_out
timeout 700 ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:front_small_kernel --launch-skip 60 --launch-count 40 -o gpurun_out/r2_front_small_src python scratch/prof_sparse.py c5 512 front > gpurun_out/r2_prof11.log 2>&1
ls -la gpurun_out/r2_front_small_src.ncu-rep
timeout 900 python bench.py > gpurun_out/r2_bench11.json 2> gpurun_out/r2_bench11.err; echo "bench rc=$?"; head -c 600 gpurun_out/r2_bench11.json; echo; tail -3 gpurun_out/r2_bench11.err
