#!/bin/bash
mkdir -p gpurun_out
run() { # name, env...
  name=$1; shift
  env "$@" timeout 200 python scratch/bench_sparse.py c5 512 front > gpurun_out/r2_sweep_$name.log 2>&1
  echo "== $name: $(grep -o "'flops': [0-9.]*\|'fronts': [0-9.]*\|'cb_doubles': [0-9.]*" gpurun_out/r2_sweep_$name.log | tr '\n' ' ')"; grep "solve (gram\|numeric\|substitutions" gpurun_out/r2_sweep_$name.log | tr '\n' ' '; echo
}
run base X=1
run tau02 THB_FRONT_TAU=0.2 THB_FRONT_MERGE_FLOPS=200000 THB_FRONT_MERGE_MAX_R=128
run tau03 THB_FRONT_TAU=0.3 THB_FRONT_MERGE_FLOPS=400000 THB_FRONT_MERGE_MAX_R=160
run tau05 THB_FRONT_TAU=0.5 THB_FRONT_MERGE_FLOPS=1000000 THB_FRONT_MERGE_MAX_R=224
run split64 THB_FRONT_SPLIT_W=64
run split128 THB_FRONT_SPLIT_W=128
run chunk128 THB_FRONT_CHUNK=128
run chunk64 THB_FRONT_CHUNK=64
run chunk256 THB_FRONT_CHUNK=256
