#!/bin/bash
# round 2, session 2: multi-GPU check of the new paths (special bonds, nve group / rmass, mixed dispersion grids)
N=${1:-2}
mkdir -p gpurun_out
out=gpurun_out/r3_mgpu_check_${N}.txt
: > $out
run() {
  echo "== N=$N $*" >> $out
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 tests/mgpu_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$" >> $out
}
run MODE=spce
run GROUP=1
run DISP=2
run DISP=3 DIFF=1
run MODE=spce GROUP=1 B200MD_P2P=0
cat $out
