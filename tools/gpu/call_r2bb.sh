mkdir -p gpurun_out
for lib in "" variants/libhtm_weave.so variants/libhtm_sqrt.so; do
  if [ -n "$lib" ]; then export HTM_B200_LIB=$PWD/$lib; else unset HTM_B200_LIB; fi
  timeout 400 python tools/lane_s_scaling.py >> gpurun_out/r2bb_s_scaling.txt 2>&1
done
cat gpurun_out/r2bb_s_scaling.txt
