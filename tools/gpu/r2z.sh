#!/bin/bash
# round 2, call Z: which CTAs (launch order) share an SM
mkdir -p gpurun_out
for args in "10000 50 100 20 5" "100000 50 60"; do
  timeout 300 python tools/gibbs_phase_trace.py $args 2>&1 | grep -v "SMs with" | tail -20
done | tee gpurun_out/r2z_cta_placement.txt
