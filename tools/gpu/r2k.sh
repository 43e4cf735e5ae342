#!/bin/bash
# round 2, call K (8 GPUs): BASELINE configs[3] over 8 and 4 GPUs, event-sharded mode C at 4 / 8, f32 shard check at 8
mkdir -p gpurun_out
nvidia-smi -L | wc -l
timeout 600 python -m pytest tests/test_select.py -m gpu -q -s > gpurun_out/r2k_pytest_select.log 2>&1; echo "select rc=$?"; grep -E "select_kernel|passed|failed" gpurun_out/r2k_pytest_select.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 8 --steps 10 --warmup 3 > gpurun_out/r2k_bench_8gpu.json 2> gpurun_out/r2k_bench_8gpu.err; echo "bench8 rc=$?"; tail -c 300 gpurun_out/r2k_bench_8gpu.err
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29522 bench.py --gpus 4 --steps 10 --warmup 3 > gpurun_out/r2k_bench_4gpu.json 2> gpurun_out/r2k_bench_4gpu.err; echo "bench4 rc=$?"
for n in 4 8; do
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 2953$n tools/gibbs_shard_perf.py 100000 50 300 2>&1 | tail -1
done | tee gpurun_out/r2k_gibbs_event_shards.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29541 tests/checks/comm_check_gibbs_f32.py 2>&1 | grep comm_check | tee gpurun_out/r2k_comm_check_f32_8gpu.txt
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29542 tests/checks/comm_check_gibbs.py 2>&1 | grep -i "oracle\|check" | tee gpurun_out/r2k_comm_check_f64_8gpu.txt
