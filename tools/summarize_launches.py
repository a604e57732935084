"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: per kernel family, launches / total us / share,
over the LAST `--last N` launches (or all).  Usage: python tools/summarize_launches.py file.csv [--last N] [--skip-last M]"""
import csv, re, sys, collections
path = sys.argv[1]
last = int(sys.argv[sys.argv.index("--last") + 1]) if "--last" in sys.argv else 0
skip_last = int(sys.argv[sys.argv.index("--skip-last") + 1]) if "--skip-last" in sys.argv else 0
rows = []
with open(path, newline="") as f:
    lines = [l for l in f if not l.startswith("==") and not l.startswith("#")]
rd = csv.reader(lines)
hdr = next(rd)
ix = {h: i for i, h in enumerate(hdr)}
for r in rd:
    if len(r) < len(hdr) or r[ix["Metric Name"]] != "gpu__time_duration.sum":
        continue
    v = float(r[ix["Metric Value"]].replace(",", ""))
    unit = r[ix["Metric Unit"]]
    us = v / 1000.0 if unit in ("ns", "nsecond") else (v * 1000.0 if unit in ("ms", "msecond") else v)
    rows.append((r[ix["Kernel Name"]], us, r[ix["Grid Size"]] if "Grid Size" in ix else ""))
if skip_last:
    rows = rows[:-skip_last]
if last:
    rows = rows[-last:]
fam = collections.OrderedDict()
for name, us, grid in rows:
    k = re.sub(r"\(.*$", "", name)
    k = re.sub(r"^void ", "", k)
    k = re.sub(r"<unnamed>::", "", k)
    k = k[:70]
    a = fam.setdefault(k, [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(v[1] for v in fam.values())
print("%d launches, %.1f us total" % (len(rows), tot))
for k, (n, us) in sorted(fam.items(), key=lambda kv: -kv[1][1]):
    print("%5d  %9.1f us  %5.1f %%  %s" % (n, us, 100 * us / tot, k))
