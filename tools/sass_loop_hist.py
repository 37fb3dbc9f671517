"""Histogram of the SASS instructions inside the hottest (largest backward-branch) loop of a kernel.
usage: python tools/sass_loop_hist.py <mangled-kernel-name> [lib]"""
import collections
import re
import subprocess
import sys

fun = sys.argv[1]
lib = sys.argv[2] if len(sys.argv) > 2 else "caribou_hi_b200/libchi.so"
txt = subprocess.run(["cuobjdump", "-sass", "-fun", fun, lib], capture_output=True, text=True).stdout
ins = []
for line in txt.splitlines():
    m = re.match(r"\s+/\*([0-9a-f]{4,5})\*/\s+(.*?);", line)
    if m:
        ins.append((int(m.group(1), 16), m.group(2).strip()))
loops = []
for addr, text in ins:
    m = re.search(r"BRA\S*\s+.*?0x([0-9a-f]+)", text)
    if m and int(m.group(1), 16) < addr:
        loops.append((addr - int(m.group(1), 16), int(m.group(1), 16), addr))
loops.sort(reverse=True)
print("loops (bytes, start, end):", [(a, hex(b), hex(c)) for a, b, c in loops[:5]])
_, lo, hi = loops[0]
body = [t for a, t in ins if lo <= a <= hi]
hist = collections.Counter()
for t in body:
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    hist[t.split()[0].split(".")[0]] += 1
fp64 = sum(v for k, v in hist.items() if k in ("DFMA", "DADD", "DMUL", "DSETP", "DMNMX"))
print("instructions in loop:", len(body), " fp64:", fp64)
for k, v in hist.most_common():
    print(f"  {k:10s} {v}")
