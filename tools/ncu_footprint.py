#!/usr/bin/env python
"""Static footprint of the executed code: `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > both.csv`
then `python tools/ncu_footprint.py both.csv <units> <src_dir>`: SASS instructions (x16 B) executed at least
thr x units times, attributed to source functions."""
import csv, sys, re, os
rows = list(csv.reader(open(sys.argv[1])))
units = float(sys.argv[2]); src_dir = sys.argv[3]
hdr = None; cur = None; line = None; text = None
recs = []  # (file, line, count)
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No": hdr = r; ix = {n: i for i, n in enumerate(hdr)}; continue
    if hdr is None or len(r) != len(hdr): continue
    if r[0] != "":
        try: line = int(r[0])
        except ValueError: line = None
        continue
    if r[2] in ("-", "...", ""): continue
    try: recs.append((cur, line, int(r[ix["Instructions Executed"]]), r[3].strip()))
    except ValueError: pass
bounds = {}
for fn in set(x[0] for x in recs):
    path = os.path.join(src_dir, fn)
    if not os.path.exists(path): continue
    marks = []
    for ln, t in enumerate(open(path).read().splitlines(), 1):
        m = re.search(r'(?:__device__|__global__|static BB_HD|^BB_HD)[^;]*?\b([A-Za-z_][A-Za-z_0-9]*)\s*\(', t)
        if m and not t.strip().startswith("//"): marks.append((ln, m.group(1)))
    bounds[fn] = marks
def owner(f, l):
    name = f
    for ln, nm in bounds.get(f, []):
        if l is not None and ln <= l: name = nm
    return name
for thr in (0.5, 0.1, 0.01, 0.0005):
    hot = [x for x in recs if x[2] >= thr * units]
    agg = {}
    for f, l, c, s in hot: agg[owner(f, l)] = agg.get(owner(f, l), 0) + 1
    print("executed >= %g x units: %d instrs = %.1f KB" % (thr, len(hot), len(hot) * 16 / 1024.0))
    print("   ", ", ".join("%s %d" % kv for kv in sorted(agg.items(), key=lambda x: -x[1])[:14]))
