#!/bin/bash
# round 2, call E: float32 mode C kernel v2 (cp.async state prefetch, uniform delta, persistent chain terms) +
# lane-kernel tuning variants
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_driver_files.py -m gpu -x -q -k "gibbs or blocked or driver or errors or switches" > gpurun_out/r2e_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2e_pytest.log
for args in "1 10 2000" "64 50 500 20 5" "1000 20 1000" "10000 50 200 20 5" "100000 50 40 20 5" "100000 50 100" "10000 20 300 20 5"; do
  timeout 300 python tools/gibbs_probe.py $args 2>&1 | tail -1
done | tee gpurun_out/r2e_gibbs_probe.txt
bash tools/gpu/r2d.sh
P="python tools/gibbs_probe.py 10000 50 20 20 5"
$P > gpurun_out/r2e_probe.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gibbs_f32_kernel -s 2 -c 1 -o gpurun_out/r2e_gibbs_f32 $P > gpurun_out/r2e_ncu.log 2>&1
echo "ncu rc=$?"
