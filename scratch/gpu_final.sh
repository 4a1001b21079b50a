#!/bin/bash
# round 2 evidence run on one B200: whole GPU suite, default bench (+ reference arm), remaining --config lines, launch list
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r2n_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2n_pytest.log
tail -6 gpurun_out/r2n_pytest.log
python bench.py > gpurun_out/r2n_bench_1gpu.json 2> gpurun_out/r2n_bench_1gpu.err; echo "bench rc=$?"; tail -c 800 gpurun_out/r2n_bench_1gpu.err
python bench.py --impl reference --steps 8 --warmup 1 > gpurun_out/r2n_bench_reference_arm.json 2> gpurun_out/r2n_bench_reference_arm.err; echo "ref arm rc=$?"
run() { name=$1; shift; python bench.py --no-cpu "$@" > gpurun_out/r2n_$name.json 2> gpurun_out/r2n_$name.err; echo "$name rc=$?"; tail -c 400 gpurun_out/r2n_$name.err; }
run spce --config spce --steps 25 --warmup 5
run spce_36k --config spce --rep 2 --steps 50 --warmup 10
run buck_big_192k --config buck_big --rep3 30 40 40 --steps 60 --warmup 10
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r2n_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-28s %9.1f M atom-steps/s  %8.3f ms/step  frac(step) %s  e2e %s parity %s" % (f[15:-5], d["value"]/1e6, d["ms_per_step"], d.get("step_roofline_frac"), d.get("e2e") and d["e2e"].get("value") and round(d["e2e"]["value"]/1e6,1), (d.get("parity") or {}).get("ok")))
        if "1gpu" in f:
            print("   cpu:", d["cpu_baseline"]["value"], d["cpu_baseline"].get("reference_loops"))
            for r in d["roofline_kernels"]: print("   %-30s %8.4f ms  %5.1f%%  frac %.3f (%s)" % (r["kernel"], r["avg_launch_ms"], 100*r["share_of_step"], r["frac"], r["bound"]))
    except Exception as e:
        print(f, "ERR", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r2n_launches.csv python bench.py --no-cpu --no-e2e --no-parity --steps 6 --warmup 3 > gpurun_out/r2n_ncu_l.log 2>&1
echo "launch list rc=$?"
python scratch/agg_launches.py gpurun_out/r2n_launches.csv 30 > gpurun_out/r2n_launches_summary.txt; cat gpurun_out/r2n_launches_summary.txt
