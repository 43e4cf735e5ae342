#!/bin/bash
# round 2, call S: lag tile 8 / 12 / 16 of the measure kernel; its GPU tests with each
mkdir -p gpurun_out
for lib in "" variants/libhtm_lag12.so variants/libhtm_lag16.so; do
  for args in "2000 50 300" "4000 20 300" "1000 50 128" "500 30 600"; do
    echo -n "lib=${lib:-default(8)}  "; HTM_B200_LIB=$lib timeout 300 python tools/measure_probe.py $args 2>&1 | tail -1
  done
  HTM_B200_LIB=$lib timeout 600 python -m pytest tests/test_measure.py -m gpu -q 2>&1 | tail -1
done | tee gpurun_out/r2s_measure_lag_tile.txt
