#!/bin/bash
# round 2, run G: register exp table (shuffle) flavour of k_pair vs the shared-memory table; parity first
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pair.py tests/test_golden.py -x -q -m gpu > gpurun_out/r2g_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2g_pytest.log
tail -4 gpurun_out/r2g_pytest.log
python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 > gpurun_out/r2g_shfl.json 2> gpurun_out/r2g_shfl.err; echo rc=$?
B200MD_PAIR_SHFL=0 python bench.py --no-cpu --no-e2e --no-parity --steps 10 --warmup 3 > gpurun_out/r2g_smem.json 2> gpurun_out/r2g_smem.err; echo rc=$?
python - <<'PY'
import json
for n in ("shfl","smem"):
    try:
        d=json.loads(open("gpurun_out/r2g_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["value"]/1e6,1), round(d["ms_per_step"],3), d["phase_ms_per_step"], (d.get("parity") or {}).get("max_rel_force_err"))
    except Exception as e: print(n, "ERR", e)
PY
KREGEX=k_pair OUT=r2_pair_shfl SKIP=2 bash scratch/gpu_ncu_k.sh
python scratch/ncu_summary.py gpurun_out/r2_pair_shfl.ncu-rep k_pair > gpurun_out/r2_pair_shfl.txt 2>&1; cat gpurun_out/r2_pair_shfl.txt | head -30
