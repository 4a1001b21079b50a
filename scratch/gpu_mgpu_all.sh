#!/bin/bash
# usage: N=4|8 bash scratch/gpu_mgpu_all.sh — the whole multi-GPU evidence run on one box
N=${N:-4}
if [ "$N" = "4" ]; then
  python -m pytest tests/test_gpu_multi.py -q -m gpu > gpurun_out/r2_pytest_multi_${N}gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2_pytest_multi_${N}gpu.log
  tail -3 gpurun_out/r2_pytest_multi_${N}gpu.log
fi
N=$N STEPS=20 bash scratch/gpu_mgpu.sh
N=$N bash scratch/gpu_mgpu2.sh
