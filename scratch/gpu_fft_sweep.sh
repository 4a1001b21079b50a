#!/bin/bash
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pppm.py -x -q 2>&1 | tail -2
for cfg in "256 8 70" "256 4 50" "256 8 50" "256 4 36" "128 4 36" "256 2 36"; do
  set -- $cfg
  export B200MD_FFT_THREADS=$1 B200MD_FFT_TBMAX=$2 B200MD_FFT_SMEM_KB=$3
  python bench.py --no-cpu --no-e2e --steps 8 --warmup 3 > gpurun_out/b_fft.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/b_fft.json'));print('$cfg', 'fft', d['phase_ms_per_step']['fft'], 'step', round(d['ms_per_step'],2))"
done
