"""Build tuning variants of libb200md.so: python scratch/variants.py name1:-DFOO=1,-DBAR=2 name2:...
Each variant recompiles only pair.cu with the extra flags and links scratch/lib_<name>.so (git-ignored *.so)."""
import os, subprocess, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import __graft_entry__ as g
pkg = g.load_package()
pkg.build()
nvcc = "/usr/local/cuda/bin/nvcc"
files = sys.argv[1].split(",")
for spec in sys.argv[2:]:
    name, _, fl = spec.partition(":")
    flags = [f for f in fl.split(",") if f]
    objs = []
    for o in sorted(os.listdir(os.path.join(pkg.HERE, "build"))):
        if not o.endswith(".o"): continue
        src = o[:-2] + ".cu"
        if src in files:
            out = os.path.join("/tmp", "var_%s_%s" % (name, o))
            cmd = [nvcc] + [f for f in pkg.NVCC_FLAGS if f != "-shared"] + pkg.extra_compile_flags() + pkg.PER_FILE_FLAGS.get(src, []) + flags + ["-c", os.path.join(pkg.CSRC, src), "-o", out]
            subprocess.run(cmd, check=True)
            objs.append(out)
        else:
            objs.append(os.path.join(pkg.HERE, "build", o))
    lib = os.path.join(pkg.ROOT, "scratch", "lib_%s.so" % name)
    subprocess.run([nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-shared", "-o", lib] + objs + pkg.extra_link_flags(), check=True)
    print("built", lib)
