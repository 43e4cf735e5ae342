#!/bin/bash
# round 2, call BO: mode C with the pair loop unrolled by 4 (default now): prefetch / ring variants, the mode C tests, the
# default bench line of the final tree
mkdir -p gpurun_out
for v in default nopf ring3; do
  unset HTM_B200_LIB HTM_GIBBS_RING
  [ $v = nopf ] && export HTM_B200_LIB=$PWD/variants/libhtm_nopf.so
  [ $v = ring3 ] && export HTM_GIBBS_RING=3
  echo "variant=$v" >> gpurun_out/r2bo_gibbs_variants.txt
  for args in "10000 50 300 20 5" "100000 50 60 20 5" "100000 50 100 4 5"; do
    timeout 200 python tools/gibbs_probe.py $args >> gpurun_out/r2bo_gibbs_variants.txt 2>&1
  done
done
unset HTM_B200_LIB HTM_GIBBS_RING
cat gpurun_out/r2bo_gibbs_variants.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "gibbs or blocked or mode_c or posterior" > gpurun_out/r2bo_pytest_subset.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2bo_pytest_subset.log
python bench.py --steps 20 --warmup 5 > gpurun_out/r2bo_bench.json 2> gpurun_out/r2bo_bench.err; echo "bench rc=$?"; tail -c 300 gpurun_out/r2bo_bench.err
