#!/bin/bash
# round 2, call X (2 GPUs): event-sharded float32 mode C after the integer limb exchange: two-GPU tests, comm check,
# time per iteration, phases
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -q -k "two_gpus" > gpurun_out/r2x_pytest_2gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2x_pytest_2gpu.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/gibbs_shard_perf.py 100000 50 300 2>&1 | tail -2 | tee gpurun_out/r2x_gibbs_shard_2gpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 tools/gibbs_phase_trace.py 100000 50 100 2>&1 | grep -v "CTA row\|event octets\|SMs with" | tail -12 | tee -a gpurun_out/r2x_gibbs_shard_2gpu.txt
