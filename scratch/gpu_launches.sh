#!/bin/bash
mkdir -p gpurun_out
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --no-cpu --no-e2e --steps 6 --warmup 3 > gpurun_out/ncu_l.log 2>&1
echo rc=$?
python scratch/agg_launches.py gpurun_out/launches.csv 16
