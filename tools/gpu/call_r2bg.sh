mkdir -p gpurun_out
for lib in variants/libhtm_old.so ""; do
  if [ -n "$lib" ]; then export HTM_B200_LIB=$PWD/$lib; else unset HTM_B200_LIB; fi
  timeout 300 python tools/lane_variant_check.py >> gpurun_out/r2bg_check.txt 2>&1
  SLOTS=1,2 timeout 300 python tools/variant_sweep.py >> gpurun_out/r2bg_sweep.txt 2>&1
done
unset HTM_B200_LIB
timeout 400 python tools/lane_s_scaling.py >> gpurun_out/r2bg_s_scaling.txt 2>&1
cat gpurun_out/r2bg_check.txt gpurun_out/r2bg_sweep.txt gpurun_out/r2bg_s_scaling.txt
timeout 900 python -m pytest tests -m gpu -x -q -k "factorised or lane or config or sharding or chunk or record or driver or posterior" > gpurun_out/r2bg_pytest_subset.log 2>&1; tail -5 gpurun_out/r2bg_pytest_subset.log
