#!/bin/bash
# usage: KREGEX=... OUT=... [SKIP=n] [BARGS='--table 1'] bash scratch/gpu_ncu_k.sh  -- ncu --set full of one launch of kernels matching KREGEX
mkdir -p gpurun_out
ncu --set full --clock-control none --import-source on --kernel-name regex:$KREGEX --launch-skip ${SKIP:-1} --launch-count ${COUNT:-1} \
  -o gpurun_out/$OUT -f python bench.py --no-cpu --no-e2e --no-parity ${BARGS:-} --steps 2 --warmup 3 > gpurun_out/ncu_k.log 2>&1
echo "ncu rc=$?"; grep -E "PROF|rror" gpurun_out/ncu_k.log | head -5
