#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pppm.py -x -q > gpurun_out/pytest_pppm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pppm.log
tail -3 gpurun_out/pytest_pppm.log
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --no-cpu --no-e2e --steps 6 --warmup 3 > gpurun_out/ncu_l.log 2>&1
echo rc=$?
python scratch/agg_launches.py gpurun_out/launches.csv 14
