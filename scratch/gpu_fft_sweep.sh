#!/bin/bash
mkdir -p gpurun_out
for cfg in "256 16 110" "256 8 110" "256 4 110" "512 16 110" "512 16 220" "256 16 220" "512 8 110" "128 4 110" "256 2 110"; do
  set -- $cfg
  export B200MD_FFT_THREADS=$1 B200MD_FFT_TBMAX=$2 B200MD_FFT_SMEM_KB=$3
  python bench.py --no-cpu --no-e2e --steps 8 --warmup 3 > gpurun_out/b_fft.json 2>/dev/null
  python -c "
import json;d=json.load(open('gpurun_out/b_fft.json'));print('$cfg', 'fft', d['phase_ms_per_step']['fft'], 'rho', d['phase_ms_per_step']['make_rho'], 'step', round(d['ms_per_step'],2))"
done
