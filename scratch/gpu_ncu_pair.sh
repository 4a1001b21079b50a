#!/bin/bash
# ncu --set full of one k_pair launch (non-EV production flavour)
mkdir -p gpurun_out
export B200MD_LIB=${B200MD_LIB_NAME:+$PWD/scratch/lib_$B200MD_LIB_NAME.so}
ncu --set full --clock-control none --import-source on --kernel-name regex:k_pair --launch-skip 3 --launch-count 1 \
  -o gpurun_out/${OUT:-prof_pair} -f python bench.py --no-cpu --no-e2e --steps 2 --warmup 3 > gpurun_out/ncu_pair.log 2>&1
echo "ncu rc=$?"; tail -3 gpurun_out/ncu_pair.log
