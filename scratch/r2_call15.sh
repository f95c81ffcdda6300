#!/bin/bash
mkdir -p gpurun_out
for pf in 1 0; do
  THB_FRONT_PREFETCH=$pf timeout 300 python scratch/bench_sparse.py c5 512 front > gpurun_out/r2_c5_512_front15_pf$pf.log 2>&1; echo "prefetch=$pf"; grep "solve (gram\|numeric" gpurun_out/r2_c5_512_front15_pf$pf.log
done
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off --csv --log-file gpurun_out/r2_c5_512_front15_launches.csv python scratch/prof_sparse.py c5 512 front > gpurun_out/r2_prof15.log 2>&1
python scratch/agg_launches.py gpurun_out/r2_c5_512_front15_launches.csv 90 > gpurun_out/r2_c5_512_front15_agg.txt 2>&1; head -9 gpurun_out/r2_c5_512_front15_agg.txt
