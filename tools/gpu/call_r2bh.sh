mkdir -p gpurun_out
for lib in variants/libhtm_old.so ""; do
  if [ -n "$lib" ]; then export HTM_B200_LIB=$PWD/$lib; else unset HTM_B200_LIB; fi
  timeout 300 python tools/lane_variant_check.py >> gpurun_out/r2bh_check.txt 2>&1
done
unset HTM_B200_LIB
SLOTS=1,2 timeout 300 python tools/variant_sweep.py >> gpurun_out/r2bh_sweep.txt 2>&1
timeout 400 python tools/lane_s_scaling.py >> gpurun_out/r2bh_s_scaling.txt 2>&1
cat gpurun_out/r2bh_check.txt gpurun_out/r2bh_sweep.txt gpurun_out/r2bh_s_scaling.txt
