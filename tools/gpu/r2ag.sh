#!/bin/bash
# round 2, call AG: radix selection with the candidates gathered into shared memory: quantile tests, detect tests + timing
mkdir -p gpurun_out
timeout 1200 python -m pytest tests/test_measure.py tests/test_gpu_posterior.py tests/test_driver_files.py -m gpu -q -s -k "detect or quantile or driver_end_to_end" > gpurun_out/r2ag_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2ag_pytest.log
grep "^detect\|\.detect" gpurun_out/r2ag_pytest.log | tee gpurun_out/r2ag_detect.txt
