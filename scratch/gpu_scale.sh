#!/bin/bash
# usage: N=2|4|8 bash scratch/gpu_scale.sh  -- the driver's launch line for the N-GPU bench
mkdir -p gpurun_out
N=${N:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29541 bench.py --gpus $N --steps 20 --warmup 5 > gpurun_out/bench_${N}gpu.json 2> gpurun_out/bench_${N}gpu.err
echo "rc=$?"; grep '^{' gpurun_out/bench_${N}gpu.json | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print(d['n_gpus'], d['value'], d['ms_per_step'], d.get('phase_ms_per_step'))"
