#!/bin/bash
# usage: N=2|4|8 [BENCH=1] bash scratch/gpu_mgpu.sh
# multi-GPU evidence run: tests/mgpu_check.py (ik, ad, + dispersion grid) with the peer-memory transposes and with the
# NCCL all-to-all, then the driver's bench line on N GPUs (slab and cube geometry, both transports)
mkdir -p gpurun_out
N=${N:-2}
OUT=gpurun_out/r2_mgpu_check_${N}.txt
: > $OUT
port=29600
tr() { port=$((port+1)); python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port $port "$@"; }
for p2p in ${P2PS:-2 0}; do
  for mode in "0 0" "1 0" "0 1"; do
    set -- $mode
    echo "== N=$N DIFF=$1 DISP=$2 B200MD_P2P=$p2p" >> $OUT
    DIFF=$1 DISP=$2 B200MD_P2P=$p2p tr tests/mgpu_check.py 2>&1 | grep -E "ranks|step|MGPU|rror|b200md:" >> $OUT
  done
done
cat $OUT
if [ "${BENCH:-1}" = "1" ]; then
  for geo in ${GEOS:-slab cube}; do
    for p2p in ${P2PS:-2 0}; do
      f=gpurun_out/r2_bench_${N}gpu_${geo}_p2p${p2p}
      B200MD_P2P=$p2p tr bench.py --gpus $N --steps ${STEPS:-20} --warmup 5 --geometry $geo > $f.json 2> $f.err
      echo "bench $geo p2p=$p2p rc=$?"; tail -c 400 $f.err
    done
  done
  python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2_bench_*gpu_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-44s N=%d %9.1f M atom-steps/s %8.3f ms/step e2e %s parity %s" % (f[11:-5], d["n_gpus"], d["value"]/1e6, d["ms_per_step"], d["e2e"] and d["e2e"].get("value") and round(d["e2e"]["value"]/1e6,1), (d.get("parity") or {}).get("max_rel_force_err")))
        print("     phases", d["phase_ms_per_step"])
    except Exception as e:
        print(f, "ERR", e)
PY
fi
