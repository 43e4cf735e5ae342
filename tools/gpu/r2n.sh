#!/bin/bash
# round 2, call N (2 GPUs): phase split of the float32 mode C kernel on 1 and 2 GPUs, per-CTA sweep spread
mkdir -p gpurun_out
for args in "10000 50 300 20 5" "100000 50 200"; do
  timeout 300 python tools/gibbs_phase_trace.py $args 2>&1 | tail -6
done | tee gpurun_out/r2n_gibbs_phases_1gpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29551 tools/gibbs_phase_trace.py 100000 50 200 2>&1 | tail -9 | tee gpurun_out/r2n_gibbs_phases_2gpu.txt
