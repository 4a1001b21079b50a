#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; python bench.py --no-cpu "$@" > gpurun_out/r2d_$name.json 2> gpurun_out/r2d_$name.err; echo "$name rc=$?"; tail -c 300 gpurun_out/r2d_$name.err; }
run spce --config spce --steps 40 --warmup 10
run spce_36k --config spce --rep 2 --steps 100 --warmup 10
run spce_table --config spce --table 1 --steps 40 --warmup 10
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2d_spce*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-28s %9.1f M atom-steps/s  %8.3f ms/step  frac(step) %s  e2e %s" % (f[15:-5], d["value"]/1e6, d["ms_per_step"], d.get("step_roofline_frac"), d.get("e2e") and d["e2e"].get("value") and round(d["e2e"]["value"]/1e6,1)), d["phase_ms_per_step"])
    except Exception as e:
        print(f, "ERR", e)
PY
