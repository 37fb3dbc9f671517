"""Builds the tracked profile summaries under profiles/ from gpurun_out/ artefacts.
usage: python tools/make_profiles.py <tag> <full.ncu-rep> <launches.csv> <bench.json>"""
import collections, csv, io, json, os, subprocess, sys

tag, rep, launches, bench = sys.argv[1:5]
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
out = os.path.join(ROOT, "profiles")
os.makedirs(out, exist_ok=True)

def run(args):
    return subprocess.run(args, capture_output=True, text=True).stdout

with open(os.path.join(out, f"{tag}_ncu_full_summary.txt"), "w") as f:
    f.write(f"# ncu --set full --clock-control none, summary of {os.path.basename(rep)}\n")
    f.write("# command: python tools/profile_target.py (configs[1] shape, 20000 draws, one launch per kernel)\n")
    f.write(run([sys.executable, os.path.join(ROOT, "tools", "ncu_summary.py"), rep]))
    for k in ("k_moments_ea", "k_moments_em", "k_moments_abs"):
        f.write("\n# dynamic opcode mix and hottest stall lines (ncu source page)\n")
        f.write(run([sys.executable, os.path.join(ROOT, "tools", "ncu_source_hist.py"), rep, k]))

rows = list(csv.reader(open(launches)))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr = rows[hi]; kn = hdr.index("Kernel Name"); mv = hdr.index("Metric Value"); gs = hdr.index("Grid Size")
agg = collections.OrderedDict()
for r in rows[hi + 1:]:
    if len(r) <= mv: continue
    key = (r[kn].split("(")[0], r[gs])
    a = agg.setdefault(key, [0, 0.0]); a[0] += 1; a[1] += float(r[mv].replace(",", ""))
tot = sum(v[1] for v in agg.values())
with open(os.path.join(out, f"{tag}_launch_list.txt"), "w") as f:
    f.write("# ncu --metrics gpu__time_duration.sum --clock-control none -c 400 python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-extras\n")
    f.write("# per-launch times are cold-cache and serialised: compare SHARES with bench.py's stage_ms\n")
    f.write(f"{'kernel':52s} {'grid':>16s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s} {'share':>7s}\n")
    for (k, g), (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write(f"{k[:52]:52s} {g:>16s} {n:8d} {t/1e3:12.1f} {t/1e3/n:10.1f} {100*t/tot:6.1f}%\n")
import shutil
shutil.copy(launches, os.path.join(out, f"{tag}_launch_list_raw.csv"))
d = json.loads(open(bench).read().strip().splitlines()[-1])
json.dump(d, open(os.path.join(out, f"{tag}_bench.json"), "w"), indent=1)
print("wrote", sorted(os.listdir(out)))
