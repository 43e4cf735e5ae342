#!/bin/bash
# round 2, call AF: smoke() with the detect stage; default bench with the detect extra
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2af_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2af_smoke.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2af_bench.json 2> gpurun_out/r2af_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2af_bench.err
