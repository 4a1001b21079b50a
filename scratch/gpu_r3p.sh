#!/bin/bash
# N-GPU bench (cubic geometry): half-spectrum transforms on / off
N=${1:-2}
mkdir -p gpurun_out
port=29700
for r in 1 0; do
  port=$((port+1))
  f=gpurun_out/r3p_bench_${N}gpu_cube_r2c$r
  B200MD_R2C=$r python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --geometry cube --no-e2e > $f.json 2> $f.err
  echo "r2c=$r rc=$?"; grep -v "OMP_NUM_THREADS\|\*\*\*\*\|^$" $f.err | tail -3
done
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r3p_bench_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-40s N=%d %9.1f M atom-steps/s %8.3f ms/step parity %s" % (f[11:-5], d["n_gpus"], d["value"]/1e6, d["ms_per_step"], (d.get("parity") or {}).get("max_rel_force_err")))
        print("     phases", d["phase_ms_per_step"])
    except Exception as e:
        print(f, "ERR", e)
PY
