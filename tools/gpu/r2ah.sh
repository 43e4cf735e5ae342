#!/bin/bash
# round 2, call AH: the measure and select file drivers, chained into the MCMC driver
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_driver_upstream.py tests/test_driver_files.py -m gpu -q > gpurun_out/r2ah_pytest.log 2>&1; echo "pytest rc=$?"; tail -25 gpurun_out/r2ah_pytest.log
