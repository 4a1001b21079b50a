#!/bin/bash
# round 2, run D: whole GPU suite, then every --config of bench.py (short runs) -> gpurun_out/r2d_*
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r2d_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2d_pytest.log
tail -8 gpurun_out/r2d_pytest.log
run() { name=$1; shift; python bench.py --no-cpu "$@" > gpurun_out/r2d_$name.json 2> gpurun_out/r2d_$name.err; echo "$name rc=$?"; tail -c 600 gpurun_out/r2d_$name.err; }
run buck --config buck --steps 40 --warmup 5
run buck_32k --config buck --rep 20 --steps 100 --warmup 10
run buck_big --config buck_big --steps 20 --warmup 5
run buck_big_192k --config buck_big --rep 0 --steps 40 --warmup 5
run buck_coul_cut --config buck_coul_cut --steps 20 --warmup 5
run buck_coul_cut_76k --config buck_coul_cut --rep 4 --steps 40 --warmup 5
run buck_coul_long_9600_1e6 --config buck_coul_long --rep 2 --acc 1e-6 --steps 40 --warmup 5
run buck_coul_long_1e6 --config buck_coul_long --acc 1e-6 --steps 10 --warmup 3
run spce_pppm_1e4 --config spce_pppm --acc 1e-4 --steps 40 --warmup 5
run spce_pppm_1e5 --config spce_pppm --acc 1e-5 --steps 40 --warmup 5
run buck_big_disp --config buck_big_disp --steps 20 --warmup 5
run table --table 1 --steps 10 --warmup 3 --no-parity
run mixed --prec mixed --steps 10 --warmup 3
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2d_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-34s %9.1f M atom-steps/s  %8.3f ms/step  frac(step) %s  parity %s" % (f[15:-5], d["value"]/1e6, d["ms_per_step"], d["step_roofline_frac"], (d.get("parity") or {}).get("ok")))
    except Exception as e:
        print(f, "ERR", e)
PY
