mkdir -p gpurun_out
for lib in "" variants/libhtm_gu1.so variants/libhtm_gu3.so variants/libhtm_gu4.so variants/libhtm_gu5.so; do
  if [ -n "$lib" ]; then export HTM_B200_LIB=$PWD/$lib; else unset HTM_B200_LIB; fi
  echo "lib=$lib" >> gpurun_out/r2bm_gibbs_unroll.txt
  for args in "10000 50 300 20 5" "100000 50 60 20 5" "100000 50 100 4 5" "10000 20 300 20 5"; do
    timeout 200 python tools/gibbs_probe.py $args >> gpurun_out/r2bm_gibbs_unroll.txt 2>&1
  done
done
cat gpurun_out/r2bm_gibbs_unroll.txt
