#!/bin/bash
# round 2, call Y: share of event octets skewed towards the CTAs an SM receives first (HTM_GIBBS_SKEW)
mkdir -p gpurun_out
for skew in 0 0.03 0.06 0.09 0.12 0.16; do
  for args in "10000 50 300 20 5" "100000 50 200" "100000 50 40 20 5" "1000 20 1000"; do
    echo -n "skew=$skew  "; HTM_GIBBS_SKEW=$skew timeout 300 python tools/gibbs_probe.py $args 2>&1 | tail -1
  done
done | tee gpurun_out/r2y_gibbs_skew.txt
for skew in 0.06 0.12; do
for args in "10000 50 300 20 5" "100000 50 100"; do
  echo "skew=$skew"; HTM_GIBBS_SKEW=$skew timeout 300 python tools/gibbs_phase_trace.py $args 2>&1 | grep -v "SMs with" | tail -16
done; done | tee -a gpurun_out/r2y_gibbs_skew.txt
