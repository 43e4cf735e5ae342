#!/bin/bash
# round 2, call AD: integer-sum range test; measure tests with the default library
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_measure.py -m gpu -q -k "integer_sums or measure" 2>&1 | tail -15
