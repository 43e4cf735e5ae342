#!/bin/bash
# round 2, call D: lane-kernel tuning variants (register caps / unroll) on the three mode B workloads
mkdir -p gpurun_out
: > gpurun_out/r2d_variants.txt
SLOTS=1,2 python tools/variant_sweep.py 2>&1 | tail -1 | tee -a gpurun_out/r2d_variants.txt
for v in variants/libhtm_*.so; do
  HTM_B200_LIB=$PWD/$v SLOTS=1,2 timeout 300 python tools/variant_sweep.py 2>&1 | tail -1 | tee -a gpurun_out/r2d_variants.txt
done
