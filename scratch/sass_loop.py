"""Opcode histogram of the hottest (innermost, largest backward-branch) loop of one SASS function.
usage: python scratch/sass_loop.py <obj> <function-substring> [print]"""
import re, subprocess, sys, collections
obj, pat = sys.argv[1], sys.argv[2]
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
funcs = out.split("Function : ")
f = [x for x in funcs if x.split("\n")[0].find(pat) >= 0][0]
ins = []
for l in f.split("\n"):
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
print(f.split("\n")[0][:100], "total instr", len(ins))
# find backward branches
loops = []
for a, t in ins:
    m = re.search(r"BRA\S*\s+(?:\S+,\s+)?(?:`\(\S+\)|0x([0-9a-f]+))", t)
    if "BRA" in t:
        m2 = re.search(r"0x([0-9a-f]+)", t.split("BRA")[1])
        if m2:
            tgt = int(m2.group(1), 16)
            if tgt < a:
                loops.append((tgt, a))
print("backward branches:", [(hex(a), hex(b), sum(1 for x, _ in ins if a <= x <= b)) for a, b in loops])
if not loops: sys.exit()
a, b = max(loops, key=lambda ab: ab[1] - ab[0]) if len(sys.argv) < 5 else loops[int(sys.argv[4])]
body = [(x, t) for x, t in ins if a <= x <= b]
h = collections.Counter()
for x, t in body:
    op = t.split()
    o = op[1] if op[0].startswith("@") else op[0]
    h[o.split(".")[0]] += 1
print("loop", hex(a), hex(b), "instr", len(body))
fp64 = sum(v for k, v in h.items() if k in ("DFMA", "DMUL", "DADD", "DSETP", "DMNMX"))
print("FP64", fp64, "other", len(body) - fp64)
print(dict(h.most_common()))
if len(sys.argv) > 3 and sys.argv[3] == "print":
    for x, t in body: print(hex(x), t)
