#!/bin/bash
python scratch/fft_sweep.py 2>&1 | tail -1
for tb in 4 16; do B200MD_FFT_TBMAX=$tb python scratch/fft_sweep.py 2>&1 | tail -1; done
for kb in 48 100 140; do B200MD_FFT_SMEM_KB=$kb python scratch/fft_sweep.py 2>&1 | tail -1; done
for th in 128 512; do B200MD_FFT_THREADS=$th python scratch/fft_sweep.py 2>&1 | tail -1; done
B200MD_FFT_SMEM_KB=100 B200MD_FFT_TBMAX=16 python scratch/fft_sweep.py 2>&1 | tail -1
B200MD_FFT_SMEM_KB=100 B200MD_FFT_THREADS=512 B200MD_FFT_TBMAX=16 python scratch/fft_sweep.py 2>&1 | tail -1
