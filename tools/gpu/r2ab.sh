#!/bin/bash
# round 2, call AB: the tree as committed: smoke(), the whole GPU suite, default bench + reference arm, launch list and
# full capture of gibbs_f32_kernel (integer sums)
mkdir -p gpurun_out
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2ab_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 gpurun_out/r2ab_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2ab_pytest.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2ab_pytest.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2ab_bench.json 2> gpurun_out/r2ab_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2ab_bench.err
python bench.py --impl reference --steps 5 --warmup 1 > gpurun_out/r2ab_bench_ref.json 2> gpurun_out/r2ab_bench_ref.err; echo "ref rc=$?"
P="python tools/gibbs_probe.py 10000 50 20 20 5"
$P > gpurun_out/r2ab_probe.txt 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/r2ab_launches_gibbs_probe.csv $P > gpurun_out/r2ab_ncu1.log 2>&1
echo "launch list rc=$?"
$P > gpurun_out/r2ab_probe2.txt 2>&1 &&
ncu --set full --clock-control none --import-source on -k regex:gibbs_f32_kernel -s 2 -c 1 -o gpurun_out/r2ab_gibbs_f32 $P > gpurun_out/r2ab_ncu2.log 2>&1
echo "full capture rc=$?"
