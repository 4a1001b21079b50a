#!/bin/bash
# round 2, run A: warp-per-atom fieldforce — parity tests, A/B bench vs the thread-per-atom kernel, launch list
mkdir -p gpurun_out
python -m pytest tests/test_gpu_pppm.py tests/test_golden.py -x -q -m gpu > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
tail -3 gpurun_out/r2a_pytest.log
python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 > gpurun_out/r2a_bench_w.json 2> gpurun_out/r2a_bench_w.err; echo rc=$?
B200MD_FF=thread python bench.py --no-cpu --no-e2e --steps 10 --warmup 3 > gpurun_out/r2a_bench_t.json 2> gpurun_out/r2a_bench_t.err; echo rc=$?
python - <<'PY'
import json
for n in ("w","t"):
    try:
        d=json.loads(open("gpurun_out/r2a_bench_%s.json"%n).read().strip().splitlines()[-1])
        print(n, d["value"]/1e6, d["ms_per_step"], d["phase_ms_per_step"])
    except Exception as e: print(n, "ERR", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2a_launches.csv python bench.py --no-cpu --no-e2e --steps 6 --warmup 3 > gpurun_out/r2a_ncu_l.log 2>&1
echo rc=$?
python scratch/agg_launches.py gpurun_out/r2a_launches.csv 24
