#!/bin/bash
mkdir -p gpurun_out
nvidia-smi -L
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > gpurun_out/r2_bench_n2.json 2> gpurun_out/r2_bench_n2.err; echo "bench n2 rc=$?"; head -c 700 gpurun_out/r2_bench_n2.json; echo; tail -4 gpurun_out/r2_bench_n2.err
timeout 600 python -m pytest tests/test_gpu_nccl_sharded.py -m gpu -q --timeout=500 -p no:cacheprovider --tb=short > gpurun_out/r2_nccl_sharded.log 2>&1; echo "nccl test rc=$?"; tail -5 gpurun_out/r2_nccl_sharded.log
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 1 --warmup 0 > gpurun_out/r2_ref_n2.json 2> gpurun_out/r2_ref_n2.err; echo "ref n2 rc=$?"; head -c 300 gpurun_out/r2_ref_n2.json
