#!/bin/bash
# round 2, call Q: station geometry from the constant bank (default build) against the expanded event rows
# (variants/libhtm_geo0.so); smoke(); the whole GPU suite
mkdir -p gpurun_out
for lib in "" variants/libhtm_geo0.so; do
  for args in "10000 50 300 20 5" "100000 50 200" "100000 50 60 20 5" "1000 20 1000" "10000 193 100 20 5"; do
    echo -n "lib=${lib:-default}  "; HTM_B200_LIB=$lib timeout 300 python tools/gibbs_probe.py $args 2>&1 | tail -1
  done
done | tee gpurun_out/r2q_gibbs_geo.txt
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2q_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 gpurun_out/r2q_smoke.log
timeout 1500 python -m pytest tests -m gpu -q > gpurun_out/r2q_pytest.log 2>&1; echo "pytest rc=$?"; tail -6 gpurun_out/r2q_pytest.log
