#!/usr/bin/env python
"""SASS evidence for the Blackwell-native kernels (CPU box, no GPU needed): `cuobjdump -sass` of the built library, one
excerpt per kernel family with the counts of the mnemonics that prove the hardware path (UTCHMMA = tcgen05.mma,
UTMALDG = TMA tensor load, UBLKCP = bulk copy, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, SYNCS = mbarrier ops) and the
instructions around the first of each.

    python scripts/sass_excerpt.py [out_dir]        # default profiles/r02
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "item_alignment_b200", "libia_b200.so")
OUT = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02")
FAMILIES = {"retrieve_tc_kernel": "sass_retrieve_tc_kernel.txt", "project_kernel": "sass_project_kernel.txt",
            "wgrad_kernel": "sass_wgrad_kernel.txt", "softmax_head_mma_kernel": "sass_softmax_head_mma_kernel.txt",
            "radix_scatter_kernel": None, "pair_kernel": None, "softmax_head_kernel": None}
MNEMONICS = ("UTCHMMA", "UTCHMMA.2CTA", "UTMALDG.2D.2CTA", "UTCBAR.2CTA", "UTCQMMA", "UTMALDG", "UTMASTG", "UBLKCP", "LDTM", "STTM", "UTCBAR", "SYNCS", "UTCATOMSWS", "LDGSTS", "LDSM", "HMMA",
             "MATCH", "REDUX", "LDG.E.128", "STG.E.128", "LDS.128", "MUFU", "IMAD", "FFMA")

sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
funcs, name, body = {}, None, []
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        if name:
            funcs[name] = body
        name, body = m.group(1), []
    elif name:
        body.append(line)
if name:
    funcs[name] = body
demangled = dict(zip(funcs, subprocess.run(["c++filt"] + list(funcs), capture_output=True, text=True).stdout.splitlines()))
os.makedirs(OUT, exist_ok=True)
index = []
for fam, fname in FAMILIES.items():
    rows = []
    for mangled, lines in funcs.items():
        dn = demangled[mangled]
        if fam not in dn:
            continue
        cnt = collections.Counter()
        for l in lines:
            for mn in MNEMONICS:
                if re.search(r"\b" + re.escape(mn), l):
                    cnt[mn] += 1
        rows.append((dn, mangled, lines, cnt))
    if not rows:
        continue
    rows.sort(key=lambda r: r[0])
    index.append((fam, rows))
    if fname is None:
        continue
    with open(os.path.join(OUT, fname), "w") as w:
        w.write(f"# cuobjdump -sass item_alignment_b200/libia_b200.so -- kernels matching '{fam}' (sm_100a)\n")
        w.write("# mnemonic counts per instantiation, then an excerpt around the first tensor-core / TMA / TMEM instructions\n\n")
        for dn, mangled, lines, cnt in rows:
            w.write(f"{dn}\n    " + "  ".join(f"{k} x{v}" for k, v in sorted(cnt.items()) if k not in ("IMAD", "FFMA")) + f"   ({len(lines)} lines)\n")
        dn, mangled, lines, cnt = rows[0]
        w.write(f"\n## excerpt: {dn}\n")
        shown = set()
        for mn in ("UTMALDG", "UBLKCP", "UTCHMMA", "UTCBAR", "LDTM", "UTMASTG", "LDGSTS", "LDSM", "HMMA"):
            for i, l in enumerate(lines):
                if re.search(r"\b" + mn, l):
                    w.write(f"\n-- first {mn} (line {i}) --\n")
                    for j in range(max(0, i - 6), min(len(lines), i + 7)):
                        if "/*" in lines[j] and re.search(r"/\*[0-9a-f]{4}\*/", lines[j]):
                            w.write(lines[j].rstrip() + "\n")
                    break
with open(os.path.join(OUT, "sass_summary.md"), "w") as w:
    w.write("# SASS mnemonic counts of the built library (`python scripts/sass_excerpt.py`, cuobjdump 12.9, sm_100a)\n\n")
    w.write("UTCHMMA = tcgen05.mma (kind::f16), UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk, LDTM = tcgen05.ld, "
            "UTCBAR = tcgen05.commit, SYNCS = mbarrier.  Full excerpts: `sass_*.txt`.\n\n")
    for fam, rows in index:
        w.write(f"## {fam} ({len(rows)} instantiations)\n\n| instantiation | " + " | ".join(MNEMONICS[:15]) + " |\n|---|" + "---|" * 15 + "\n")
        for dn, mangled, lines, cnt in rows[:40]:
            short = re.sub(r"^void ia::", "", dn).split("(")[0]
            w.write(f"| `{short}` | " + " | ".join(str(cnt.get(m, 0)) for m in MNEMONICS[:15]) + " |\n")
        w.write("\n")
print("wrote", OUT)
