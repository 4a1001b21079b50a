#!/bin/bash
# usage: N=2|4|8 bash scratch/gpu_mgpu2.sh — config 5 (buck/long/coul/long + pppm/disp) and strong scaling on N GPUs
mkdir -p gpurun_out
N=${N:-2}
port=29700
tr() { port=$((port+1)); python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
run() { name=$1; shift; tr bench.py --gpus $N "$@" > gpurun_out/r2_${name}_${N}gpu.json 2> gpurun_out/r2_${name}_${N}gpu.err; echo "$name rc=$?"; grep -v "OMP_NUM_THREADS\|\*\*\*\*\|^$" gpurun_out/r2_${name}_${N}gpu.err | tail -5; }
run disp_cube --config buck_big_disp --geometry cube --steps 20 --warmup 5 --no-parity
run strong --scaling strong --rep 15 --steps 20 --warmup 5 --no-parity
python - <<'PY'
import json, glob, os
N=os.environ.get("N","2")
for f in sorted(glob.glob("gpurun_out/r2_*_%sgpu.json" % N)):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-40s N=%d atoms %s %9.1f M atom-steps/s %8.3f ms/step e2e %s" % (os.path.basename(f)[3:-5], d["n_gpus"], d["config"]["workload"].split("(")[1].split(")")[0], d["value"]/1e6, d["ms_per_step"], d["e2e"] and d["e2e"].get("value") and round(d["e2e"]["value"]/1e6,1)))
        print("     phases", d["phase_ms_per_step"])
    except Exception as e:
        print(f, "ERR", e)
PY
