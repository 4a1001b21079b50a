#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pppm.py tests/test_golden.py tests/test_host.py -x -q -m gpu > gpurun_out/pytest_pppm.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pppm.log
tail -3 gpurun_out/pytest_pppm.log
run() { python bench.py --no-cpu --no-e2e --steps 20 --warmup 5 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d.get('phase_ms_per_step'))"; }
run
B200MD_PREC=mixed run
