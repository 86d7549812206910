#!/usr/bin/env python
"""Per-source-line summary of an ncu report: `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > both.csv`
then `python tools/ncu_lines.py both.csv [top]`.  Prints instructions executed and stall samples per source line."""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
hdr = None; cur = None; out = []
for r in rows:
    if len(r) >= 2 and r[0] == "File Path": cur = r[1].split("/")[-1]; continue
    if len(r) >= 2 and r[0] == "Line No": hdr = r; continue
    if hdr is None or len(r) != len(hdr) or r[0] == "": continue
    d = dict(zip(range(len(hdr)), r))
    try:
        ix = {n: i for i, n in enumerate(hdr)}
        out.append((cur, int(r[0]), r[1].strip()[:90], int(r[ix["Instructions Executed"]]), int(r[ix["# Samples"]]),
                    {n: int(r[i]) for n, i in ix.items() if n.startswith("stall_") and "Not Issued" not in n}))
    except ValueError:
        pass
ti = sum(o[3] for o in out); ts = sum(o[4] for o in out)
print("total instr %d samples %d" % (ti, ts))
agg = {}
for o in out:
    for k, v in o[5].items(): agg[k] = agg.get(k, 0) + v
print("stalls:", {k: "%.1f%%" % (100.0 * v / max(ts, 1)) for k, v in sorted(agg.items(), key=lambda x: -x[1])[:8]})
print("--- by instructions")
for o in sorted(out, key=lambda o: -o[3])[:top]:
    print("%5.1f%% i %5.1f%% s  %s:%d  %s" % (100.0 * o[3] / ti, 100.0 * o[4] / ts, o[0], o[1], o[2]))
print("--- by samples")
for o in sorted(out, key=lambda o: -o[4])[:top]:
    st = sorted(o[5].items(), key=lambda x: -x[1])[:2]
    print("%5.1f%% s %5.1f%% i  %s:%d  %s   %s" % (100.0 * o[4] / ts, 100.0 * o[3] / ti, o[0], o[1], o[2], st))

# optional: aggregate by function (line ranges found by scanning the source for "__device__" / "__global__")
if len(sys.argv) > 3:
    import re, os
    src_dir = sys.argv[3]
    bounds = {}
    for fn in set(o[0] for o in out):
        path = os.path.join(src_dir, fn)
        if not os.path.exists(path): continue
        marks = []
        lines = open(path).read().splitlines()
        for ln, text in enumerate(lines, 1):
            m = re.search(r'(?:__device__|__global__|static BB_HD|^BB_HD)[^;]*?\b([A-Za-z_][A-Za-z_0-9]*)\s*\(', text)
            if m and not text.strip().startswith("//"): marks.append((ln, m.group(1)))
        bounds[fn] = marks
    agg = {}
    for o in out:
        name = o[0]
        for ln, nm in bounds.get(o[0], []):
            if ln <= o[1]: name = o[0] + ":" + nm
        a = agg.setdefault(name, [0, 0]); a[0] += o[3]; a[1] += o[4]
    print("--- by function")
    for k, v in sorted(agg.items(), key=lambda x: -x[1][0])[:30]:
        print("%5.1f%% i %5.1f%% s  %s" % (100.0 * v[0] / ti, 100.0 * v[1] / ts, k))
