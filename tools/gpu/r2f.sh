#!/bin/bash
# round 2, call F: float32 mode C kernel variants (register budget / unroll / prefetch) + new tests (summary, posterior)
mkdir -p gpurun_out
: > gpurun_out/r2f_gibbs_variants.txt
for lib in default variants/libhtm_g*.so; do
  for args in "10000 50 200 20 5" "100000 50 60"; do
    if [ "$lib" = default ]; then r=$(timeout 300 python tools/gibbs_probe.py $args 2>&1 | tail -1); else r=$(HTM_B200_LIB=$PWD/$lib timeout 300 python tools/gibbs_probe.py $args 2>&1 | tail -1); fi
    echo "$lib  $r" | tee -a gpurun_out/r2f_gibbs_variants.txt
  done
done
timeout 1500 python -m pytest tests -m gpu -q -x -k "summary or posterior or quantile or first_iteration or shared_parameter_ratio or many_joint" > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?"; tail -8 gpurun_out/r2f_pytest.log
