#!/bin/bash
# round 2, run B: ncu --set full of both fieldforce kernels, the z/poisson pass and the FFT passes (summaries for profiles/)
mkdir -p gpurun_out
KREGEX=k_fieldforce_w OUT=r2_ff_w SKIP=2 bash scratch/gpu_ncu_k.sh
B200MD_FF=thread KREGEX=k_fieldforce OUT=r2_ff_t SKIP=2 bash scratch/gpu_ncu_k.sh
KREGEX='k_fft' OUT=r2_fft SKIP=6 COUNT=6 bash scratch/gpu_ncu_k.sh
for r in r2_ff_w r2_ff_t; do python scratch/ncu_summary.py gpurun_out/$r.ncu-rep k_fieldforce > gpurun_out/$r.txt 2>&1; done
ls -la gpurun_out/*.ncu-rep
