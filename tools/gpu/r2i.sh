#!/bin/bash
# round 2, call I (2 GPUs): multi-GPU tests again (itemised f32 check, late shard, time-out), wide tempering groups
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -k "wide_groups or argument_and_state or fp64 or abi" > gpurun_out/r2i_pytest_1gpu.log 2>&1; echo "pytest 1gpu rc=$?"; tail -5 gpurun_out/r2i_pytest_1gpu.log
timeout 1500 python -m pytest tests -m gpu -q -k "two_gpus" > gpurun_out/r2i_pytest_2gpu.log 2>&1; echo "pytest 2gpu rc=$?"; tail -8 gpurun_out/r2i_pytest_2gpu.log; grep -n "comm_check" gpurun_out/r2i_pytest_2gpu.log | head
