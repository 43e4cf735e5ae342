#!/bin/bash
# round 2, call R: the measure stage (htm_measure.cu) against its oracle; DFMA rate at 2 000 x 50 x 300
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_measure.py -m gpu -q -s > gpurun_out/r2r_pytest.log 2>&1; echo "pytest rc=$?"; tail -12 gpurun_out/r2r_pytest.log
grep "measure_kernel" gpurun_out/r2r_pytest.log > gpurun_out/r2r_measure.txt
