#!/bin/bash
# round 2, call B: the new float32 mode C kernel (tests first, then timing), the refactored host path
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q -k "gibbs or blocked or driver or abi or smoke or client" > gpurun_out/r2b_pytest_gibbs.log 2>&1; echo "pytest gibbs rc=$?"; tail -15 gpurun_out/r2b_pytest_gibbs.log
timeout 900 python -m pytest tests -m gpu -x -q -k "not gibbs and not blocked" > gpurun_out/r2b_pytest_rest.log 2>&1; echo "pytest rest rc=$?"; tail -5 gpurun_out/r2b_pytest_rest.log
for args in "1 10 2000" "1000 20 1000" "10000 50 200 20 5" "100000 50 40 20 5" "100000 50 100" "10000 20 300 20 5"; do
  timeout 300 python tools/gibbs_probe.py $args 2>&1 | tail -1
done | tee gpurun_out/r2b_gibbs_probe.txt
timeout 600 python bench.py --steps 3 --warmup 3 --no-cpu > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"; tail -c 400 gpurun_out/r2b_bench.err
