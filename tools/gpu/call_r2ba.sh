set -x
mkdir -p gpurun_out
for lib in "" variants/libhtm_weave.so variants/libhtm_sqrt.so variants/libhtm_weavesqrt.so; do
  if [ -n "$lib" ]; then export HTM_B200_LIB=$PWD/$lib; else unset HTM_B200_LIB; fi
  timeout 300 python tools/lane_variant_check.py >> gpurun_out/r2ba_check.txt 2>&1
  SLOTS=1,2 timeout 300 python tools/variant_sweep.py >> gpurun_out/r2ba_sweep.txt 2>&1
done
cat gpurun_out/r2ba_check.txt gpurun_out/r2ba_sweep.txt
