#!/bin/bash
mkdir -p gpurun_out
for T in 512 256 128; do
  timeout 400 ncu --set full --clock-control none --import-source on --profile-from-start off --kernel-name-base demangled -k "regex:front_small_kernel<\\(int\\)$T>" --launch-skip 6 --launch-count 1 -o gpurun_out/r2_fs_src_$T python scratch/prof_sparse.py c5 512 front > gpurun_out/r2_prof12_$T.log 2>&1
done
ls -la gpurun_out/r2_fs_src_*.ncu-rep; tail -3 gpurun_out/r2_prof12_512.log
