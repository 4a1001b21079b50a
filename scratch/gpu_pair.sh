#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_pair.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pair.log
tail -5 gpurun_out/pytest_pair.log
for lib in "" $VARIANTS; do
  if [ -n "$lib" ]; then export B200MD_LIB=$PWD/scratch/lib_$lib.so; fi
  python bench.py --no-cpu --no-e2e --steps 10 > gpurun_out/bench_pair_$lib.json 2> gpurun_out/bench_pair_$lib.err; echo "bench $lib rc=$?"
  python -c "
import json;d=json.load(open('gpurun_out/bench_pair_$lib.json'));print('$lib', d['value'],d['ms_per_step'],d['roofline']['frac'],d['phase_ms_per_step'])"
done
