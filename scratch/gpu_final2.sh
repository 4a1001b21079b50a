#!/bin/bash
# round 2 final evidence run on one B200: whole GPU suite, default bench (+ reference arm), the PPPM-heavy --config lines
# (their FFT phase changed with the half-spectrum transforms), launch list, ncu --set full of the top / new kernels
mkdir -p gpurun_out
python -m pytest tests -q -m gpu > gpurun_out/r3f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r3f_pytest.log
tail -5 gpurun_out/r3f_pytest.log
python bench.py > gpurun_out/r3f_bench_1gpu.json 2> gpurun_out/r3f_bench_1gpu.err; echo "bench rc=$?"; tail -c 500 gpurun_out/r3f_bench_1gpu.err
python bench.py --impl reference --steps 8 --warmup 1 > gpurun_out/r3f_bench_reference_arm.json 2> gpurun_out/r3f_bench_reference_arm.err; echo "ref arm rc=$?"
run() { name=$1; shift; python bench.py --no-cpu "$@" > gpurun_out/r3f_cfg_$name.json 2> gpurun_out/r3f_cfg_$name.err; echo "$name rc=$?"; tail -c 300 gpurun_out/r3f_cfg_$name.err; }
run buck_coul_long_1e6 --config buck_coul_long --acc 1e-6 --steps 10 --warmup 3
run spce_pppm_1e4 --config spce_pppm --acc 1e-4 --steps 40 --warmup 5
run spce_pppm_1e5 --config spce_pppm --acc 1e-5 --steps 40 --warmup 5
run buck_big_disp --config buck_big_disp --steps 20 --warmup 5
run spce --config spce --steps 25 --warmup 5
python - <<'PY'
import json, glob
for f in sorted(glob.glob("gpurun_out/r3f_*.json")):
    try:
        d=json.loads(open(f).read().strip().splitlines()[-1])
        print("%-34s %9.1f M atom-steps/s  %8.3f ms/step  frac(step) %s  e2e %s parity %s" % (f[15:-5], d["value"]/1e6, d["ms_per_step"], d.get("step_roofline_frac"), d.get("e2e") and d["e2e"].get("value") and round(d["e2e"]["value"]/1e6,1), (d.get("parity") or {}).get("ok")))
        if "1gpu" in f:
            print("   cpu:", d["cpu_baseline"]["value"])
            for r in d["roofline_kernels"]: print("   %-30s %8.4f ms  %5.1f%%  frac %.3f (%s)" % (r["kernel"], r["avg_launch_ms"], 100*r["share_of_step"], r["frac"], r["bound"]))
    except Exception as e:
        print(f, "ERR", e)
PY
ncu --metrics gpu__time_duration.sum --clock-control none -c 700 --csv --log-file gpurun_out/r3f_launches.csv python bench.py --no-cpu --no-e2e --no-parity --steps 6 --warmup 3 > gpurun_out/r3f_ncu_l.log 2>&1
echo "launch list rc=$?"
python scratch/agg_launches.py gpurun_out/r3f_launches.csv 30 > gpurun_out/r3f_launches_summary.txt; head -24 gpurun_out/r3f_launches_summary.txt
for k in k_pair k_fft_z_poisson k_fft_x_r2c k_fft_x_c2r; do
  ncu --set full --clock-control none --import-source on --kernel-name regex:$k --launch-skip 3 --launch-count 1 -o gpurun_out/r3f_$k -f python bench.py --no-cpu --no-e2e --no-parity --steps 2 --warmup 3 > gpurun_out/r3f_ncu_$k.log 2>&1
  echo "ncu $k rc=$?"
done
ls -la gpurun_out/r3f_*.ncu-rep
