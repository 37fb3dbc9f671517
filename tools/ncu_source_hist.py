"""Dynamic opcode histogram + hottest stall lines from an ncu source page.
usage: python tools/ncu_source_hist.py report.ncu-rep [kernel-substring]"""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; pat = sys.argv[2] if len(sys.argv) > 2 else ''
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks = txt.split('"Kernel Name",')
seen = set()
for blk in blocks[1:]:
    lines = blk.splitlines()
    name = lines[0].strip('",')
    if pat not in name or name in seen: continue
    seen.add(name)
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rows[0]
    iS, iE, iSt = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    hist = collections.Counter(); tot = 0; stall = []
    for r in rows[1:]:
        if len(r) <= iE: continue
        src = r[iS].strip(); n = int(r[iE] or 0)
        toks = src.split()
        if not toks: continue
        op = toks[1] if toks[0].startswith('@') else toks[0]
        op = op.split('.')[0].rstrip(';')
        hist[op] += n; tot += n
        stall.append((int(r[iSt] or 0), src))
    print("====", name, "total warp-instr", tot)
    for op, n in hist.most_common(22):
        print(f"  {op:10s} {n:12d} {100.0 * n / tot:5.1f}%")
    print("  -- top stall lines")
    for s, src in sorted(stall, reverse=True)[:14]:  # noqa
        print(f"  {s:6d}  {src[:90]}")
