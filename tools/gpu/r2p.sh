#!/bin/bash
# round 2, call P: TMA ring depth 2 / 3 / 4 in the float32 mode C kernel; mode C tests
mkdir -p gpurun_out
for ring in 2 3 4; do
  for args in "10000 50 300 20 5" "100000 50 200" "100000 50 60 20 5" "1000 20 1000"; do
    echo -n "ring=$ring  "; HTM_GIBBS_RING=$ring timeout 300 python tools/gibbs_probe.py $args 2>&1 | tail -1
  done
done | tee gpurun_out/r2p_gibbs_ring.txt
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_gpu_posterior.py -m gpu -q -k "gibbs or blocked or config0 or quantiles" > gpurun_out/r2p_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2p_pytest.log
