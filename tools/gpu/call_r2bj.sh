mkdir -p gpurun_out
for rows in 0 auto; do
  if [ "$rows" = auto ]; then unset HTM_GIBBS_ROWS; else export HTM_GIBBS_ROWS=$rows; fi
  echo "HTM_GIBBS_ROWS=$rows" >> gpurun_out/r2bj_gibbs_rows_auto.txt
  for args in "10000 50 300 20 5" "100000 50 60 20 5" "100000 50 100 4 5" "10000 50 300 8 5" "10000 50 200 16 5" "10000 50 200 16 8" "10000 50 100 30 20" "1000 20 1000 4 5"; do
    timeout 200 python tools/gibbs_probe.py $args >> gpurun_out/r2bj_gibbs_rows_auto.txt 2>&1
  done
done
cat gpurun_out/r2bj_gibbs_rows_auto.txt
unset HTM_GIBBS_ROWS
timeout 900 python -m pytest tests -m gpu -x -q -k "gibbs or blocked or mode_c or posterior" > gpurun_out/r2bj_pytest_subset.log 2>&1; tail -5 gpurun_out/r2bj_pytest_subset.log
