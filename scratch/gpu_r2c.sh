#!/bin/bash
# round 2, run C: whole GPU suite + the new bench line (default workload) + reference arm
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -5 gpurun_out/r2c_pytest.log
python bench.py > gpurun_out/r2c_bench.json 2> gpurun_out/r2c_bench.err; echo "bench rc=$?"
tail -c 3000 gpurun_out/r2c_bench.err
python - <<'PY'
import json
d=json.loads(open("gpurun_out/r2c_bench.json").read().strip().splitlines()[-1])
print("value %.1f M  ms/step %.3f  e2e %.1f M  step_roofline_frac %s" % (d["value"]/1e6, d["ms_per_step"], d["e2e"]["value"]/1e6, d["step_roofline_frac"]))
print("parity", d["parity"])
for r in d["roofline_kernels"]: print("  %-30s %8.4f ms  %5.1f%% of step  frac %.3f (%s)" % (r["kernel"], r["avg_launch_ms"], 100*r["share_of_step"], r["frac"], r["bound"]))
print("cpu", d["cpu_baseline"]["value"] if d["cpu_baseline"] else None)
PY
