#!/usr/bin/env python
"""tools/sass_lines.py <cubin> <kernel substring> [pattern]: instructions (and those matching `pattern`, default LDL|STL)
of one kernel per source line, from `nvdisasm -g` (the cubin comes from `cuobjdump -xelf all build/bb_nvN.o`)."""
import re, subprocess, sys, collections
cubin, kern = sys.argv[1], sys.argv[2]
pat = re.compile(sys.argv[3] if len(sys.argv) > 3 else r"LDL|STL")
txt = subprocess.run(["nvdisasm", "-g", cubin], capture_output=True, text=True).stdout.splitlines()
inside = False; cur = None; cnt = collections.Counter(); tot = collections.Counter()
for l in txt:
    if l.startswith("//---") and ".text." in l:
        inside = kern in l
        continue
    if not inside: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m: cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    if re.search(r"/\*[0-9a-f]{4,}\*/", l):
        tot[cur] += 1
        if pat.search(l): cnt[cur] += 1
print("instructions %d, matching %d" % (sum(tot.values()), sum(cnt.values())))
for k, v in cnt.most_common(40): print(k, v, "of", tot[k])
