#!/usr/bin/env python3
"""Instruction share by source function from an ncu source-page csv:  ncu -i rep --page source --csv --print-source cuda,sass > f.csv"""
import bisect, collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1])))
cur = hdr = None
agg = collections.defaultdict(lambda: [0, 0, 0])
for r in rows:
    if not r: continue
    if r[0] in ("File Path", "File Name"): cur = r[1].split("/")[-1]; continue
    if r[0] == "Line No": hdr = r; continue
    if hdr is None or "Instructions Executed" not in hdr: continue
    try:
        line = int(r[0]); inst = int(r[hdr.index("Instructions Executed")]); th = int(r[hdr.index("Thread Instructions Executed")]); sa = int(r[hdr.index("# Samples")])
    except Exception: continue
    a = agg[(cur, line)]; a[0] += inst; a[1] += th; a[2] += sa
tot = sum(a[0] for a in agg.values()) or 1; tott = sum(a[1] for a in agg.values()) or 1; tots = sum(a[2] for a in agg.values()) or 1
print("warp inst", tot, "thread inst", tott, "samples", tots)
def marks(path):
    out = []
    for i, l in enumerate(open(path).read().split("\n"), 1):
        m = re.match(r"^(?:  )?(?:PG_HDN?|PG_MEMBER|PG_PHILOX_ATTR|PG_HOSTDEV|__global__|static|template <[^>]*>\s*PG_HD)\b.*?(\w+)\(", l)
        if m: out.append((i, m.group(1)))
    return out
for f in sorted({k[0] for k in agg}):
    path = "pgtg_b200/csrc/" + f
    try: mk = marks(path)
    except Exception: continue
    lines = [m[0] for m in mk]
    by = collections.defaultdict(lambda: [0, 0, 0])
    for (ff, ln), a in agg.items():
        if ff != f: continue
        k = bisect.bisect_right(lines, ln) - 1
        name = mk[k][1] if k >= 0 else "?"
        b = by[name]; b[0] += a[0]; b[1] += a[1]; b[2] += a[2]
    print("==", f, "%.1f%%" % (100 * sum(b[0] for b in by.values()) / tot))
    for name, b in sorted(by.items(), key=lambda kv: -kv[1][0])[:14]:
        if b[0] * 300 < tot: continue
        print("   %-30s warp-inst %5.1f%%  thread-inst %5.1f%%  lanes %4.1f  samples %5.1f%%" % (name, 100 * b[0] / tot, 100 * b[1] / tott, b[1] / max(b[0], 1), 100 * b[2] / tots))
