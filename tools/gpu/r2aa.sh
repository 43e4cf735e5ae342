#!/bin/bash
# round 2, call AA (8 GPUs): event-sharded float32 mode C after the integer limb exchange: shard check at 8, time per
# iteration at 8 / 4, phases at 8, bench --gpus 8 (selfcheck + mode C extra)
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/checks/comm_check_gibbs_f32.py 2>&1 | grep comm_check | tee gpurun_out/r2aa_comm_check_f32_8gpu.txt
for n in 8 4; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n tools/gibbs_shard_perf.py 100000 50 300 2>&1 | tail -1
done | tee gpurun_out/r2aa_gibbs_event_shards.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29513 tools/gibbs_phase_trace.py 100000 50 100 2>&1 | grep -v "CTA row\|event octets\|SMs with\|SM \|SM of" | tail -12 | tee -a gpurun_out/r2aa_gibbs_event_shards.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2aa_bench_8gpu.json 2> gpurun_out/r2aa_bench_8gpu.err; echo "bench8 rc=$?"; tail -c 300 gpurun_out/r2aa_bench_8gpu.err
