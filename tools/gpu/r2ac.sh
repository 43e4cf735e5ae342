#!/bin/bash
# round 2, call AC: measure kernel with the unit queue, 4 / 6 / 10 units per warp
mkdir -p gpurun_out
for lib in variants/libhtm_mu4.so "" variants/libhtm_mu10.so; do
for args in "2000 50 300" "4000 20 300" "1000 50 128" "500 30 600" "4000 10 300"; do
  echo -n "lib=${lib:-default(6)}  "; HTM_B200_LIB=$lib timeout 300 python tools/measure_probe.py $args 2>&1 | tail -1
done; done | tee gpurun_out/r2ac_measure_units.txt
timeout 600 python -m pytest tests/test_measure.py -m gpu -q 2>&1 | tail -3
