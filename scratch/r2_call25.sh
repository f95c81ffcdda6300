#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 2 --steps 3 --warmup 3 --no-c2 > gpurun_out/r2_bench_n2_25.json 2> gpurun_out/r2_bench_n2_25.err; echo "bench n2 rc=$?"; grep '^{' gpurun_out/r2_bench_n2_25.json | head -c 700; echo; tail -3 gpurun_out/r2_bench_n2_25.err
