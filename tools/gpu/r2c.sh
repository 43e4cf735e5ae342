#!/bin/bash
# round 2, call C: full GPU suite with the new float32 mode C kernel + posterior tests, ncu of gibbs_f32_kernel
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2c_pytest.log
P="python tools/gibbs_probe.py 10000 50 20 20 5"
$P > gpurun_out/r2c_probe.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gibbs_f32_kernel -s 2 -c 1 -o gpurun_out/r2c_gibbs_f32 $P > gpurun_out/r2c_ncu.log 2>&1
echo "ncu rc=$?"; cat gpurun_out/r2c_probe.txt
