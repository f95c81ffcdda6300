#!/bin/bash
mkdir -p gpurun_out
for st in 0 5120; do
  THB_SOLVE_STAGE=$st timeout 300 python scratch/bench_sparse.py c5 512 front > gpurun_out/r2_c5_512_front18_st$st.log 2>&1; echo "stage=$st"; grep "lm_it\|solve (gram\|substitutions" gpurun_out/r2_c5_512_front18_st$st.log | cut -c1-160
done
timeout 300 python scratch/prof_lm_step.py 512 > gpurun_out/r2_lm_step18.log 2>&1; echo "prof rc=$?"; head -48 gpurun_out/r2_lm_step18.log | cut -c1-170
NCU="ncu --set full --import-source on --clock-control none --profile-from-start off --kernel-name-base demangled"
timeout 300 $NCU -k "regex:front_forward_kernel<\(int\)256>" -s 12 -c 1 -f -o gpurun_out/r2_fwd256 python scratch/prof_sparse.py c5 512 front > gpurun_out/r2_ncu18a.log 2>&1; echo "ncu fwd rc=$?"
timeout 300 $NCU -k "regex:front_backward_kernel<\(int\)256>" -s 3 -c 1 -f -o gpurun_out/r2_bwd256 python scratch/prof_sparse.py c5 512 front > gpurun_out/r2_ncu18b.log 2>&1; echo "ncu bwd rc=$?"
timeout 300 $NCU -k "regex:front_small_kernel<\(int\)1024>" -s 7 -c 1 -f -o gpurun_out/r2_fs1024 python scratch/prof_sparse.py c5 512 front > gpurun_out/r2_ncu18c.log 2>&1; echo "ncu fs1024 rc=$?"
ls -la gpurun_out/*.ncu-rep
