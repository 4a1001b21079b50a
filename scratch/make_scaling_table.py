"""profiles/r02_bench_*gpu_*.json, r02_strong_*gpu.json, r02_disp_cube_*gpu.json -> profiles/r02_scaling.md"""
import glob
import json
import os
import re


def load(f):
    return json.loads(open(f).read().strip().splitlines()[-1])


one = load("profiles/r02_configs_buck_big_disp.json")
base = {"bench": None, "disp": one["value"]}
rows = []
for f in sorted(glob.glob("profiles/r02_bench_*gpu_*.json")) + sorted(glob.glob("profiles/r02_strong_*gpu.json")) + \
        sorted(glob.glob("profiles/r02_disp_cube_*gpu.json")):
    d = load(f)
    rows.append((os.path.basename(f)[4:-5], d))
b1 = [d for n, d in rows if n == "bench_1gpu"]
with open("profiles/r02_scaling.md", "w") as fh:
    fh.write("# Multi-GPU runs of round 2 (one 8 x B200 box, `python -m torch.distributed.run ... bench.py --gpus N`)\n\n")
    fh.write("Device-timed steps (max over ranks), atoms resident; `e2e` = b200md_step_host_ids (ids, x, f to pinned host "
             "memory every step).  Transposes: p2p2 = strided copy-engine copies into peer memory (default), p2p1 = stores "
             "from a scatter kernel, p2p0 = NCCL all-to-all with pack / unpack kernels.  `parity` = max force error of the "
             "614 k-atom sample decomposed over the same N ranks against the CPU oracle.  Efficiency is NOT computed here "
             "(the driver does that from its own runs); the single-GPU lines are in r02_configs.md / r02_bench_1gpu.json.\n\n")
    fh.write("| run | N | atoms | atom-steps/s | ms/step | e2e atom-steps/s | parity (force) | phases ms/step |\n|---|---|---|---|---|---|---|---|\n")
    for n, d in rows:
        atoms = re.search(r"\((\d+) atoms\)", d["config"]["workload"]).group(1)
        par = (d.get("parity") or {}).get("max_rel_force_err")
        e2e = d.get("e2e") or {}
        ph = ", ".join("%s %.2f" % (k, v) for k, v in d["phase_ms_per_step"].items() if v >= 0.005)
        fh.write("| %s | %d | %s | %.1f M | %.3f | %s | %s | %s |\n" % (
            n, d["n_gpus"], atoms, d["value"] / 1e6, d["ms_per_step"],
            ("%.1f M" % (e2e["value"] / 1e6)) if e2e.get("value") else "-", ("%.1e" % par) if par is not None else "-", ph))
print(open("profiles/r02_scaling.md").read())
