#!/usr/bin/env python
"""Stall samples / executed instructions of one kernel of an .ncu-rep aggregated per CUDA source line (CPU box).

    ncu -i rep --page source --csv > src.csv;  nvdisasm -g -c file.cubin > dis.txt
    python scripts/ncu_source_lines.py src.csv dis.txt '<mangled kernel name>' [csrc dir]
"""
import collections, csv, re, sys
src_csv, dis_txt, mangled = sys.argv[1:4]
csrc = sys.argv[4] if len(sys.argv) > 4 else "item_alignment_b200/csrc/"
dis = open(dis_txt).read().split('\n')
start = [i for i, l in enumerate(dis) if l.startswith('.text.' + mangled + ':')][0]
line_of_off, cur = {}, None
for l in dis[start + 1:]:
    if l.startswith('//---------------------'):
        break
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m = re.match(r'\s+/\*([0-9a-f]{4,})\*/', l)
    if m:
        line_of_off[int(m.group(1), 16)] = cur
rows = list(csv.reader(open(src_csv)))
hi = [i for i, r in enumerate(rows) if r and r[0] == "Address"][0]
hdr = rows[hi]
data = [r for r in rows[hi + 1:] if len(r) >= len(hdr)]
idx = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
base = int(data[0][idx['Address']], 16)
agg = collections.defaultdict(lambda: [0, 0, collections.Counter()])
for r in data:
    ln = line_of_off.get(int(r[idx['Address']], 16) - base)
    try:
        agg[ln][0] += int(r[idx['# Samples']]); agg[ln][1] += int(r[idx['Instructions Executed']])
        for h in stall_cols:
            if r[idx[h]] not in ('', '0'):
                agg[ln][2][h[6:]] += int(r[idx[h]])
    except ValueError:
        pass
T = sum(v[0] for v in agg.values()) or 1
I = sum(v[1] for v in agg.values()) or 1
files = {}
print(f"total samples {T}, warp instructions {I}\nsamples% instr%  file:line  top stalls | source")
for ln, (s, i, st) in sorted(agg.items(), key=lambda x: -x[1][0])[:int(sys.argv[5]) if len(sys.argv) > 5 else 40]:
    text = ''
    if ln:
        if ln[0] not in files:
            try:
                files[ln[0]] = open(csrc + ln[0]).read().split('\n')
            except OSError:
                files[ln[0]] = []
        if ln[1] - 1 < len(files[ln[0]]):
            text = files[ln[0]][ln[1] - 1].strip()[:90]
    print(f"{100 * s / T:5.1f} {100 * i / I:5.1f}  {ln[0] if ln else None}:{ln[1] if ln else ''}  {dict(st.most_common(3))} | {text}")
