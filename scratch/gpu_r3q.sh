#!/bin/bash
# 4-GPU bench line (cubic geometry) with the half-spectrum transforms
N=${1:-4}
mkdir -p gpurun_out
f=gpurun_out/r3q_bench_${N}gpu_cube
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29801 bench.py --gpus $N --steps 20 --warmup 5 --geometry cube > $f.json 2> $f.err
echo "rc=$?"; grep -v "OMP_NUM_THREADS\|\*\*\*\*\|^$" $f.err | tail -3
python - <<PY
import json
d=json.loads(open("$f.json").read().strip().splitlines()[-1])
print("N=%d %9.1f M atom-steps/s %8.3f ms/step e2e %s parity %s" % (d["n_gpus"], d["value"]/1e6, d["ms_per_step"], d["e2e"] and round(d["e2e"]["value"]/1e6,1), (d.get("parity") or {}).get("max_rel_force_err")))
print("     phases", d["phase_ms_per_step"])
PY
