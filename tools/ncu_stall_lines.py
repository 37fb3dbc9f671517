"""Main-loop listing of a kernel from an ncu source page: per SASS line the executed count, total
stall samples and the dominant stall reasons.  usage: python tools/ncu_stall_lines.py rep kernel-substr [min_exec]"""
import csv, io, subprocess, sys
rep, pat = sys.argv[1], sys.argv[2]
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
blocks = txt.split('"Kernel Name",')
for blk in blocks[1:]:
    lines = blk.splitlines()
    if pat not in lines[0]:
        continue
    rows = list(csv.reader(io.StringIO("\n".join(lines[1:]))))
    hdr = rows[0]
    iS, iE, iA = hdr.index("Source"), hdr.index("Instructions Executed"), hdr.index("Warp Stall Sampling (All Samples)")
    reasons = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
    mx = max(int(r[iE] or 0) for r in rows[1:] if len(r) > iE)
    thr = float(sys.argv[3]) if len(sys.argv) > 3 else 0.5
    tot = {}
    for r in rows[1:]:
        if len(r) <= iE or int(r[iE] or 0) < thr * mx:
            continue
        rs = sorted(((int(r[i] or 0), h[6:]) for i, h in reasons), reverse=True)[:2]
        for i, h in reasons:
            tot[h] = tot.get(h, 0) + int(r[i] or 0)
        print(f"{int(r[iA] or 0):5d} {r[iS][:60]:60s} " + " ".join(f"{h}={n}" for n, h in rs if n))
    print({k: v for k, v in sorted(tot.items(), key=lambda kv: -kv[1]) if v})
    break
