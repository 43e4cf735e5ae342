#!/bin/bash
# round 2, call BL (2 GPUs): the multi-GPU tests, bench --gpus 2 (configs[3] split in two) and mode C event shards on the
# tree with the phased lane kernel and the mode C row choice
mkdir -p gpurun_out
nvidia-smi -L | head -3
timeout 600 python -m pytest tests -m gpu -q -k "two_gpus" > gpurun_out/r2bl_pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2bl_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2bl_bench_2gpu.json 2> gpurun_out/r2bl_bench_2gpu.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2bl_bench_2gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/gibbs_shard_perf.py 100000 50 300 2>&1 | tail -2 | tee gpurun_out/r2bl_gibbs_shard_2gpu.txt
