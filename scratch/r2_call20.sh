#!/bin/bash
mkdir -p gpurun_out
run() { tag=$1; shift; env "$@" timeout 300 python scratch/bench_sparse.py c5 512 front > gpurun_out/r2_c5_512_front20_$tag.log 2>&1; echo "$tag"; grep "lm_it\|numeric\|substitutions" gpurun_out/r2_c5_512_front20_$tag.log | cut -c1-130; }
run base THB_X=0
run t1_256 THB_FRONT_T1=256
run t0_128 THB_FRONT_T0=128
run t0_128_t1_256 THB_FRONT_T0=128 THB_FRONT_T1=256
