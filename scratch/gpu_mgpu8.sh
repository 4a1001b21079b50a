#!/bin/bash
# N=8 evidence run: correctness check, bench (slab: copy-engine transposes and NCCL; cube: copy-engine), config 5 at 32 M, strong scaling
mkdir -p gpurun_out
export N=8
N=8 P2PS="2 0" BENCH=0 bash scratch/gpu_mgpu.sh
port=29800
tr() { port=$((port+1)); python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $port "$@"; }
run() { f=gpurun_out/r2_bench_8gpu_$1; shift; "$@" > $f.json 2> $f.err; echo "$f rc=$?"; grep -v "OMP_NUM_THREADS\|\*\*\*\*\|^$" $f.err | tail -3; }
B200MD_P2P=2 run slab_p2p2 tr bench.py --gpus 8 --steps 20 --warmup 5 --geometry slab
B200MD_P2P=0 run slab_p2p0 tr bench.py --gpus 8 --steps 20 --warmup 5 --geometry slab
B200MD_P2P=2 run cube_p2p2 tr bench.py --gpus 8 --steps 20 --warmup 5 --geometry cube
N=8 bash scratch/gpu_mgpu2.sh
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_8gpu_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-44s N=%d %9.1f M atom-steps/s %8.3f ms/step e2e %s parity %s" % (f[11:-5], d["n_gpus"], d["value"]/1e6, d["ms_per_step"], d["e2e"] and d["e2e"].get("value") and round(d["e2e"]["value"]/1e6,1), (d.get("parity") or {}).get("max_rel_force_err")))
        print("     phases", d["phase_ms_per_step"])
    except Exception as e:
        print(f, "ERR", e)
PY
