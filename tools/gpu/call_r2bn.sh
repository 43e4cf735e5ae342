mkdir -p gpurun_out
for lib in variants/libhtm_gu4.so variants/libhtm_gu6.so variants/libhtm_gu8.so variants/libhtm_gu12.so; do
  export HTM_B200_LIB=$PWD/$lib
  echo "lib=$lib" >> gpurun_out/r2bm_gibbs_unroll.txt
  for args in "10000 50 300 20 5" "100000 50 60 20 5" "100000 50 100 4 5" "10000 20 300 20 5" "10000 51 300 20 5"; do
    timeout 200 python tools/gibbs_probe.py $args >> gpurun_out/r2bm_gibbs_unroll.txt 2>&1
  done
done
for lib in "" variants/libhtm_lu3.so variants/libhtm_lu4.so; do
  if [ -n "$lib" ]; then export HTM_B200_LIB=$PWD/$lib; else unset HTM_B200_LIB; fi
  timeout 300 python tools/lane_variant_check.py >> gpurun_out/r2bn_lane_unroll.txt 2>&1
  SLOTS=1,2 timeout 300 python tools/variant_sweep.py >> gpurun_out/r2bn_lane_unroll.txt 2>&1
done
tail -22 gpurun_out/r2bm_gibbs_unroll.txt; cat gpurun_out/r2bn_lane_unroll.txt
