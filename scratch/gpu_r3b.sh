#!/bin/bash
# half-spectrum (r2c) PPPM vs the complex-to-complex passes on the bench workload
mkdir -p gpurun_out
for r in 1 0; do
  B200MD_R2C=$r python bench.py --no-cpu --no-e2e --steps 20 --warmup 5 > gpurun_out/r3b_bench_r2c$r.json 2> gpurun_out/r3b_bench_r2c$r.err; echo "r2c=$r rc=$?"
done
python - <<'PY'
import json
for r in (1, 0):
    d=json.loads(open("gpurun_out/r3b_bench_r2c%d.json" % r).read().strip().splitlines()[-1])
    print("r2c=%d" % r, round(d["value"]/1e6,1), d["ms_per_step"], d["phase_ms_per_step"], (d.get("parity") or {}).get("ok"), (d.get("parity") or {}).get("ekspace_rel"))
    for k in d["roofline_kernels"]:
        if "fft" in k["kernel"]: print("   %-30s %8.4f ms x %d" % (k["kernel"], k["avg_launch_ms"], k["launches"]))
PY
