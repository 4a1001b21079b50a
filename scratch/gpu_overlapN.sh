#!/bin/bash
mkdir -p gpurun_out
N=${N:-4}
B200MD_OVERLAP=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29551 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu --no-e2e 2>&1 | grep '^{' | tail -1 | python -c "
import sys, json
d = json.loads(sys.stdin.readline()); print(d['n_gpus'], d['value'], d['ms_per_step'], d.get('phase_ms_per_step'))"
