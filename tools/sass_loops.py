"""Loops of a kernel's SASS with instruction histograms (development aid).
usage: python tools/sass_loops.py <object-or-lib> <mangled-kernel-name> [min_instructions]"""
import collections, re, subprocess, sys
obj, fun = sys.argv[1], sys.argv[2]
nmin = int(sys.argv[3]) if len(sys.argv) > 3 else 100
txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, obj], capture_output=True, text=True).stdout
ins = []
for line in txt.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
loops = []
for addr, text in ins:
    m = re.search(r"BRA\S*\s+.*?0x([0-9a-f]+)", text)
    if m and int(m.group(1), 16) < addr:
        loops.append((int(m.group(1), 16), addr))
print("total instructions", len(ins))
for lo, hi in sorted(loops):
    body = [t for a, t in ins if lo <= a <= hi]
    if len(body) < nmin:
        continue
    inner = sum(1 for l2, h2 in loops if l2 >= lo and h2 <= hi) - 1
    hist = collections.Counter(re.sub(r"^@!?U?P\d+\s+", "", t).split()[0].split(".")[0] for t in body)
    fp64 = sum(v for k, v in hist.items() if k in ("DFMA", "DADD", "DMUL", "DSETP"))
    print(f"{lo:#x}-{hi:#x} n={len(body)} fp64={fp64} inner_loops={inner} " + " ".join(f"{k}:{v}" for k, v in hist.most_common(14)))
