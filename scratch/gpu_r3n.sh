#!/bin/bash
# multi-GPU check of the half-spectrum transforms through all three transposes
N=${1:-2}
mkdir -p gpurun_out
out=gpurun_out/r3n_mgpu_check_${N}.txt
: > $out
run() {
  echo "== N=$N $*" >> $out
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29521 tests/mgpu_check.py 2>&1 | grep -v "^\*\*\*\|OMP_NUM_THREADS\|^$\|NCCL version" >> $out
}
run DIFF=0
#run DIFF=1
#run DISP=1 B200MD_P2P=0
run DISP=2 B200MD_P2P=1
#run DIFF=0 B200MD_R2C=0
cat $out
