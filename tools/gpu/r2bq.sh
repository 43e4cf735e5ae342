#!/bin/bash
# round 2, call BQ: launch list and full capture of gibbs_f32_kernel of the final tree (5 CTA rows, pair loop unrolled by 4,
# incremental ring indices) at the sample shape, 10 000 events x 50 stations x 100 joint chains
mkdir -p gpurun_out
P="python tools/gibbs_probe.py 10000 50 20 20 5"
$P > gpurun_out/r2bq_probe.txt 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2bq_launches_gibbs_probe.csv $P > gpurun_out/r2bq_ncu1.log 2>&1
echo "launch list rc=$?"
$P > gpurun_out/r2bq_probe2.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gibbs_f32_kernel -s 2 -c 1 -o gpurun_out/r2bq_gibbs_f32 $P > gpurun_out/r2bq_ncu2.log 2>&1
echo "full capture rc=$?"
cat gpurun_out/r2bq_probe.txt
