#!/bin/bash
# round 2, call W: float32 mode C with exact integer sums (atomics into two limbs) in the place of per-CTA partials
mkdir -p gpurun_out
for args in "10000 50 300 20 5" "100000 50 200" "100000 50 60 20 5" "1000 20 1000" "1 10 2000"; do
  timeout 300 python tools/gibbs_probe.py $args 2>&1 | tail -1
done | tee gpurun_out/r2w_gibbs_exact_sums.txt
for args in "10000 50 300 20 5" "100000 50 100"; do
  timeout 300 python tools/gibbs_phase_trace.py $args 2>&1 | grep -v "CTA row\|event octets\|SMs with" | tail -12
done | tee -a gpurun_out/r2w_gibbs_exact_sums.txt
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_gpu_posterior.py -m gpu -q -k "gibbs or blocked or config0 or quantiles" > gpurun_out/r2w_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2w_pytest.log
