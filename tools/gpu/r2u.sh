#!/bin/bash
# round 2, call U: ncu capture of measure_kernel (296 windows x 50 x 300 = two full waves)
mkdir -p gpurun_out
timeout 300 python tools/measure_probe.py 296 50 300 2>&1 | tail -1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:measure_kernel -c 1 -o gpurun_out/r2u_measure python tools/measure_probe.py 296 50 300 > gpurun_out/r2u_ncu.log 2>&1; echo "ncu rc=$?"; tail -2 gpurun_out/r2u_ncu.log
