#!/bin/bash
# round 2, call BK: GPU suite, smoke(), default bench + reference arm of the tree with the phased lane kernel and the
# mode C row choice; launch list and one full capture of the lane kernel
mkdir -p gpurun_out
nvidia-smi --query-gpu=index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active --format=csv -lms 500 > gpurun_out/r2bk_clocks.csv &
SMI=$!
python -m pytest tests -m gpu -x -q > gpurun_out/r2bk_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2bk_pytest_gpu.log
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2bk_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2bk_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2bk_bench.json 2> gpurun_out/r2bk_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2bk_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2bk_bench_ref.json 2> gpurun_out/r2bk_bench_ref.err; echo "ref rc=$?"
kill $SMI
SHORT="python bench.py --steps 2 --warmup 1 --no-cpu --no-extra --iters 2000 --interval 100"
$SHORT > gpurun_out/r2bk_short.json 2> gpurun_out/r2bk_short.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2bk_launches_bench_short.csv $SHORT > gpurun_out/r2bk_ncu1.log 2>&1
echo "launch list rc=$?"
$SHORT > gpurun_out/r2bk_short2.json 2> gpurun_out/r2bk_short2.err &&
ncu --set full --clock-control none --import-source on -k regex:fact_lane_kernel -s 2 -c 1 -o gpurun_out/r2bk_lane_configs2 $SHORT > gpurun_out/r2bk_ncu2.log 2>&1
echo "full capture rc=$?"
ls -la gpurun_out | tail -12
