#!/bin/bash
# round 2, call M: whole GPU suite (timed) with the balanced chain rows; mode C probes
mkdir -p gpurun_out
S0=$(date +%s)
timeout 2400 python -m pytest tests -m gpu -q --durations=12 > gpurun_out/r2m_pytest.log 2>&1; echo "pytest rc=$? in $(( $(date +%s) - S0 )) s"; tail -22 gpurun_out/r2m_pytest.log
for args in "1 10 2000" "1000 20 1000" "10000 50 300 20 5" "100000 50 100 20 5" "100000 50 200" "10000 20 300 20 5"; do
  timeout 300 python tools/gibbs_probe.py $args 2>&1 | tail -1
done | tee gpurun_out/r2m_gibbs_probe.txt
