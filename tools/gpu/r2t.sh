#!/bin/bash
# round 2, call T: measure kernel with the register sliding window (16 lags x 16 steps)
mkdir -p gpurun_out
for args in "2000 50 300" "4000 20 300" "1000 50 128" "500 30 600" "2000 50 301"; do
  timeout 300 python tools/measure_probe.py $args 2>&1 | tail -1
done | tee gpurun_out/r2t_measure.txt
timeout 600 python -m pytest tests/test_measure.py -m gpu -q 2>&1 | tail -3
