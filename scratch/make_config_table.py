"""gpurun_out/<prefix>*.json (bench.py --config runs) -> profiles/<out>.md + the JSON lines copied to profiles/.
usage: python scratch/make_config_table.py r2d_ r02_configs"""
import glob
import json
import os
import shutil
import sys

prefix, out = sys.argv[1], sys.argv[2]
rows = []
for f in sorted(glob.glob("gpurun_out/%s*.json" % prefix)):
    try:
        d = json.loads(open(f).read().strip().splitlines()[-1])
    except (ValueError, IndexError):
        continue
    name = os.path.basename(f)[len(prefix):-5]
    shutil.copy(f, "profiles/%s_%s.json" % (out, name))
    rows.append((name, d))
with open("profiles/%s.md" % out, "w") as fh:
    fh.write("# bench.py --config runs, one B200 (%s)\n\n" % out)
    fh.write("Each row is one `bench.py` JSON line (kept beside this file as `%s_<name>.json`): device-timed steps, atoms "
             "resident.\n`frac(step)` = sum of the kernels' roofline-ideal times / measured step; per-kernel fractions are "
             "in the JSON (`roofline_kernels`).\n\n" % out)
    fh.write("| run | workload | atom-steps/s | ms/step | phases (ms/step) | k_pair frac | frac(step) | parity |\n")
    fh.write("|---|---|---|---|---|---|---|---|\n")
    for name, d in rows:
        ph = ", ".join("%s %.2f" % (k, v) for k, v in d["phase_ms_per_step"].items() if v >= 0.005)
        kp = [r for r in d["roofline_kernels"] if r["kernel"] == "k_pair"]
        par = d.get("parity") or {}
        ptxt = "n/a" if not par or "skipped" in par else ("force %.1e, E %.1e, set %s" % (
            par.get("max_rel_force_err", float("nan")), par.get("epair_rel", float("nan")), par.get("pair_set_equal")))
        fh.write("| %s | %s | %.1f M | %.3f | %s | %s | %s | %s |\n" % (
            name, d["config"]["workload"], d["value"] / 1e6, d["ms_per_step"], ph,
            ("%.2f (%s)" % (kp[0]["frac"], kp[0]["bound"])) if kp else "-", d["step_roofline_frac"], ptxt))
print(open("profiles/%s.md" % out).read())
