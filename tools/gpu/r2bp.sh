#!/bin/bash
# round 2, call BP: mode C with the ring indices advanced by increments (no 64-bit divisions per visit); whole GPU suite
mkdir -p gpurun_out
for args in "10000 50 300 20 5" "100000 50 60 20 5" "100000 50 100 4 5" "10000 20 300 20 5" "1 10 4000 4 5"; do
  timeout 200 python tools/gibbs_probe.py $args >> gpurun_out/r2bp_gibbs_ring_indices.txt 2>&1
done
cat gpurun_out/r2bp_gibbs_ring_indices.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r2bp_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2bp_pytest_gpu.log
