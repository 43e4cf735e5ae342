#!/bin/bash
# round 2, call J (2 GPUs): flag-based exchange in the float32 mode C kernel -- tests on 1 and 2 GPUs, timings
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q -k "gibbs or blocked or wide_groups or driver or full_size or config0 or quantiles" > gpurun_out/r2j_pytest_1gpu.log 2>&1; echo "pytest 1gpu rc=$?"; tail -6 gpurun_out/r2j_pytest_1gpu.log
timeout 1500 python -m pytest tests -m gpu -q -k "two_gpus" > gpurun_out/r2j_pytest_2gpu.log 2>&1; echo "pytest 2gpu rc=$?"; tail -4 gpurun_out/r2j_pytest_2gpu.log
for args in "1 10 2000" "64 50 500 20 5" "1000 20 1000" "10000 50 200 20 5" "100000 50 100"; do
  timeout 300 python tools/gibbs_probe.py $args 2>&1 | tail -1
done | tee gpurun_out/r2j_gibbs_probe.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 tools/gibbs_shard_perf.py 100000 50 300 2>&1 | tail -1 | tee gpurun_out/r2j_gibbs_shard_2gpu.txt
