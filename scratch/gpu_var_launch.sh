#!/bin/bash
mkdir -p gpurun_out
for lib in "" $VARIANTS; do
  if [ -n "$lib" ]; then export B200MD_LIB=$PWD/scratch/lib_$lib.so; fi
  ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches_$lib.csv python bench.py --no-cpu --no-e2e --steps 6 --warmup 3 > gpurun_out/ncu_l.log 2>&1
  echo "== $lib"; python scratch/agg_launches.py gpurun_out/launches_$lib.csv 12 | grep -E "$GREP"
done
