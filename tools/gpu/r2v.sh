#!/bin/bash
# round 2, call V: which CTAs of the float32 mode C sweep are late (row, octet count, SM sharing)
mkdir -p gpurun_out
for args in "10000 50 300 20 5" "100000 50 100"; do
  timeout 300 python tools/gibbs_phase_trace.py $args 2>&1 | tail -28
done | tee gpurun_out/r2v_gibbs_cta_spread.txt
