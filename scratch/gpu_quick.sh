#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pppm.py tests/test_gpu_pair.py tests/test_golden.py -x -q -m gpu > gpurun_out/pytest_q.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_q.log
tail -3 gpurun_out/pytest_q.log
python bench.py --no-cpu --no-e2e --steps 20 --warmup 5 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d.get('phase_ms_per_step'))"
ncu --metrics gpu__time_duration.sum --clock-control none -c 500 --csv --log-file gpurun_out/launches.csv python bench.py --no-cpu --no-e2e --steps 5 --warmup 3 > gpurun_out/ncu_l.log 2>&1
python scratch/agg_launches.py gpurun_out/launches.csv 14 | grep -E "nb_|rho_"
