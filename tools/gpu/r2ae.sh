#!/bin/bash
# round 2, call AE: detect stage (correlation functions of all windows, thresholds by radix selection, detection) against its
# oracle; one day of 50 stations; measure tests again (the kernel is shared)
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_measure.py -m gpu -q -s > gpurun_out/r2ae_pytest.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/r2ae_pytest.log
grep "^detect\|measure_kernel" gpurun_out/r2ae_pytest.log > gpurun_out/r2ae_detect.txt
timeout 300 python tools/measure_probe.py 2000 50 300 | tail -1
