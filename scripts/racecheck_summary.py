#!/usr/bin/env python
"""Condense a compute-sanitizer racecheck log (which repeats one 12-line record per byte address) into one line per
(kind, reader site, writer site) with a count.  Runs on the GPU box right after the sanitizer, so only the summary travels.

    compute-sanitizer --tool racecheck --racecheck-report all --print-limit 1000000 python scripts/sanitize_small.py > log 2>&1
    python scripts/racecheck_summary.py log > summary.txt
"""
import collections
import re
import sys

recs = collections.Counter()
kind = rd = wr = None
summary = []
for line in open(sys.argv[1], errors="replace"):
    m = re.search(r"(Error|Warning): (Potential )?(\w+) hazard detected at (__shared__|__global__)?", line)
    if m:
        kind = f"{m.group(1)} {m.group(3)} {m.group(4) or ''}".strip()
        rd = wr = None
        continue
    if kind:
        m = re.search(r"(Read|Write) Thread .* at (.*?)\+0x[0-9a-f]+ in (\S+)", line)
        if m:
            site = f"{m.group(2).split('(')[0][:60]} @ {m.group(3)}"
            if m.group(1) == "Read" and rd is None:
                rd = site
            elif wr is None:
                wr = site
            elif rd is None:
                rd = site
        if "Current Value" in line:
            recs[(kind, rd, wr)] += 1
            kind = None
    if "RACECHECK SUMMARY" in line or "ERROR SUMMARY" in line:
        summary.append(line.strip())
print("# racecheck records grouped by (kind, first access site, second access site); count = byte addresses reported")
for (k, r, w), c in recs.most_common():
    print(f"{c:8d}  {k}\n            first : {r}\n            second: {w}")
for s in summary:
    print(s)
