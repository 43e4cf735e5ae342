#!/bin/bash
# round 2, call O: long-run reproducibility test, final default bench + reference arm, launch list of a mode C probe,
# final full capture of gibbs_f32_kernel
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py -m gpu -q -k "long_run or chunked or many_joint or first_iteration" > gpurun_out/r2o_pytest.log 2>&1; echo "pytest rc=$?"; tail -4 gpurun_out/r2o_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2o_bench.json 2> gpurun_out/r2o_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2o_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2o_bench_ref.json 2> gpurun_out/r2o_bench_ref.err; echo "ref rc=$?"
P="python tools/gibbs_probe.py 10000 50 20 20 5"
$P > gpurun_out/r2o_probe.txt 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2o_launches_gibbs_probe.csv $P > gpurun_out/r2o_ncu1.log 2>&1
echo "launch list rc=$?"
$P > gpurun_out/r2o_probe2.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gibbs_f32_kernel -s 2 -c 1 -o gpurun_out/r2o_gibbs_f32 $P > gpurun_out/r2o_ncu2.log 2>&1
echo "full capture rc=$?"
for args in "10000 50 300 20 5" "100000 50 200"; do timeout 300 python tools/gibbs_probe.py $args 2>&1 | tail -1; done
