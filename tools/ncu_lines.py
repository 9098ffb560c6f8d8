#!/usr/bin/env python
"""Per-source-line instruction / stall-sample shares of one kernel from an ncu report.
ncu's CSV export of the CUDA-source view carries no metrics, so this joins the SASS view
(`ncu -i rep --page source --csv`) with `nvdisasm -g -c` line info of the same cubin by
instruction order.  usage: ncu_lines.py <sass.csv> <nvdisasm.txt> <kernel-symbol-substring> [top]"""
import csv, re, sys
from collections import defaultdict

sass_csv, dis, sym = sys.argv[1:4]
top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
lines = open(dis).read().split("\n")
start = next(i for i, l in enumerate(lines) if l.startswith(".text.") and sym in l)
end = next((i for i in range(start + 1, len(lines)) if lines[i].startswith(".text.") or lines[i].startswith(".section")), len(lines))
cur, ins = None, []
for l in lines[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1), int(m.group(2)))
        continue
    m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", l)
    if m:
        ins.append((m.group(2).strip(), cur))
rows = list(csv.reader(open(sass_csv)))
h = next(i for i, r in enumerate(rows) if "Source" in r and "# Samples" in r)
hdr = rows[h]; ci = {n: k for k, n in enumerate(hdr)}
sass = [r for r in rows[h + 1:] if len(r) == len(hdr)]
assert len(sass) == len(ins), (len(sass), len(ins))
by = defaultdict(lambda: [0, 0]); tot = ts = 0
for r, (txt, cur) in zip(sass, ins):
    n = int(r[ci["Instructions Executed"]]); s = int(r[ci["# Samples"]])
    by[cur][0] += n; by[cur][1] += s; tot += n; ts += s
print(f"warp instructions {tot}, samples {ts}")
cache = {}
for k, v in sorted(by.items(), key=lambda kv: -kv[1][0])[:top]:
    f, ln = k if k else ("?", 0)
    if f not in cache:
        try: cache[f] = open(f).read().split("\n")
        except OSError: cache[f] = []
    text = cache[f][ln - 1].strip()[:100] if 0 < ln <= len(cache[f]) else ""
    print(f"{v[0]/tot*100:5.1f}% inst {v[1]/ts*100:5.1f}% smp {f.split('/')[-1]}:{ln}: {text}")
