#!/bin/bash
mkdir -p gpurun_out
timeout 300 python scratch/bench_sparse.py c5 512 front > gpurun_out/r2_c5_512_front17_nograph.log 2>&1; grep "lm_it_per_s" gpurun_out/r2_c5_512_front17_nograph.log
timeout 300 python scratch/bench_sparse.py c5 512 front graph > gpurun_out/r2_c5_512_front17_graph.log 2>&1; grep "lm_it_per_s" gpurun_out/r2_c5_512_front17_graph.log; tail -3 gpurun_out/r2_c5_512_front17_graph.log
