#!/bin/bash
mkdir -p gpurun_out
run() { python bench.py --no-cpu --no-e2e --steps 20 --warmup 5 2>&1 | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d.get('phase_ms_per_step'))"; }
echo "== HEAD"; B200MD_LIB=scratch/lib_head.so run
echo "== new MAXRADIX=5"; B200MD_FFT_MAXRADIX=5 run
echo "== HEAD"; B200MD_LIB=scratch/lib_head.so run
