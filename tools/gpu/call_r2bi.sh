mkdir -p gpurun_out
for rows in 0 5 6 7 9 13; do
  export HTM_GIBBS_ROWS=$rows
  echo "HTM_GIBBS_ROWS=$rows" >> gpurun_out/r2bi_gibbs_rows.txt
  timeout 200 python tools/gibbs_probe.py 10000 50 300 20 5 >> gpurun_out/r2bi_gibbs_rows.txt 2>&1
  timeout 200 python tools/gibbs_probe.py 100000 50 60 20 5 >> gpurun_out/r2bi_gibbs_rows.txt 2>&1
  timeout 200 python tools/gibbs_probe.py 10000 20 300 20 5 >> gpurun_out/r2bi_gibbs_rows.txt 2>&1
done
cat gpurun_out/r2bi_gibbs_rows.txt
