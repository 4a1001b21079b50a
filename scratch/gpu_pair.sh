#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pair.py -x -q > gpurun_out/pytest_pair.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pair.log
tail -5 gpurun_out/pytest_pair.log
python bench.py --no-cpu --no-e2e --steps 10 > gpurun_out/bench_pair.json 2> gpurun_out/bench_pair.err; echo "bench rc=$?"
python -c "
import json;d=json.load(open('gpurun_out/bench_pair.json'));print(d['value'],d['ms_per_step'],d['roofline']['frac'],d['phase_ms_per_step'])"
