#!/bin/bash
# round 2, call G: whole GPU suite + default bench + reference arm
mkdir -p gpurun_out
timeout 2400 python -m pytest tests -m gpu -q > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2g_pytest.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2g_smoke.log
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r2g_bench.json 2> gpurun_out/r2g_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2g_bench.err
