#!/bin/bash
# round 2, call L: whole GPU suite (timed), select kernel v2, per-iteration phase split of the float32 mode C kernel
mkdir -p gpurun_out
/usr/bin/time -v timeout 2400 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r2l_pytest.log 2> gpurun_out/r2l_pytest.time; echo "pytest rc=$?"; tail -22 gpurun_out/r2l_pytest.log; grep -E "Elapsed|Maximum resident" gpurun_out/r2l_pytest.time
python -m pytest tests/test_select.py -m gpu -q -s 2>&1 | grep -E "select_kernel|passed|failed" | tee gpurun_out/r2l_select.txt
for args in "1 10 2000" "1000 20 1000" "10000 50 300 20 5" "100000 50 200"; do
  timeout 300 python tools/gibbs_phase_trace.py $args 2>&1 | tail -6
done | tee gpurun_out/r2l_gibbs_phases_1gpu.txt
