#!/bin/bash
# round 2, run F: fast Coulomb-table flavour — parity (pair + golden tests), A/B bench vs the generic flavour, ncu
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pair.py tests/test_golden.py tests/test_host.py -x -q -m gpu > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -5 gpurun_out/r2f_pytest.log
python bench.py --no-cpu --no-e2e --table 1 --steps 10 --warmup 3 > gpurun_out/r2f_table_fast.json 2> gpurun_out/r2f_table_fast.err; echo rc=$?
B200MD_TABFAST=0 python bench.py --no-cpu --no-e2e --no-parity --table 1 --steps 10 --warmup 3 > gpurun_out/r2f_table_generic.json 2> gpurun_out/r2f_table_generic.err; echo rc=$?
python bench.py --no-cpu --no-e2e --no-parity --steps 10 --warmup 3 > gpurun_out/r2f_analytic.json 2> gpurun_out/r2f_analytic.err; echo rc=$?
python - <<'PY'
import json
for n in ("table_fast","table_generic","analytic"):
    try:
        d=json.loads(open("gpurun_out/r2f_%s.json"%n).read().strip().splitlines()[-1])
        print(n, round(d["value"]/1e6,1), round(d["ms_per_step"],3), d["phase_ms_per_step"], (d.get("parity") or {}).get("max_rel_force_err"))
    except Exception as e: print(n, "ERR", e)
PY
BARGS='--table 1' KREGEX=k_pair OUT=r2_pair_table SKIP=2 bash scratch/gpu_ncu_k.sh

python scratch/ncu_summary.py gpurun_out/r2_pair_table.ncu-rep k_pair > gpurun_out/r2_pair_table.txt 2>&1; cat gpurun_out/r2_pair_table.txt | head -30
KREGEX=k_rho_tiles OUT=r2_rho_tiles SKIP=2 bash scratch/gpu_ncu_k.sh
python scratch/ncu_summary.py gpurun_out/r2_rho_tiles.ncu-rep k_rho_tiles > gpurun_out/r2_rho_tiles.txt 2>&1; head -12 gpurun_out/r2_rho_tiles.txt
