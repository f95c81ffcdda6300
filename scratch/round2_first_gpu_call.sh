#!/bin/bash
# Round 2, first GPU call:   gpurun --timeout 2400 -- 'bash scratch/round2_first_gpu_call.sh'
# Runs everything that was written after the round-1 GPU budget was spent (tests/test_gpu_zz_first_run.py), then times the new sparse
# layouts on the C5 shape.  Every step under its own timeout; logs in gpurun_out/.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_zz_first_run.py tests/first_run_kernels.py -m gpu -q -rxX --timeout=400 -p no:cacheprovider > gpurun_out/r02_pending.log 2>&1
echo "pending rc=$?" >> gpurun_out/r02_pending.log
tail -5 gpurun_out/r02_pending.log
for B in 128 512; do
  for lay in lane lane_root lane_tiled lane_tiled_root; do
    timeout 240 python scratch/bench_sparse.py c5 $B $lay > gpurun_out/r02_c5_${B}_${lay}.log 2>&1
    timeout 240 python scratch/bench_sparse.py c5 $B $lay supernodal > gpurun_out/r02_c5_${B}_${lay}_supernodal.log 2>&1
  done
done
grep -H "lm_it_per_s\|numeric\|substitutions ms" gpurun_out/r02_c5_*.log | tail -60
