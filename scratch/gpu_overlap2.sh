#!/bin/bash
mkdir -p gpurun_out
N=${N:-2}
fmt() { grep '^{' | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print(d['value'], d['ms_per_step'], d.get('phase_ms_per_step'))"; }
run() { python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $1 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-e2e 2>&1 | fmt; }
echo "== 1 GPU serial"; python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e 2>&1 | fmt
echo "== 1 GPU overlap"; B200MD_OVERLAP=1 python bench.py --steps 20 --warmup 5 --no-cpu --no-e2e 2>&1 | fmt
echo "== $N GPU serial"; run 29511
echo "== $N GPU overlap"; B200MD_OVERLAP=1 run 29512
B200MD_OVERLAP=1 python -m pytest tests/test_gpu_multi.py tests/test_gpu_pppm.py -x -q 2>&1 | tail -2
