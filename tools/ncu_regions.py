#!/usr/bin/env python
"""Share of executed instructions / stall samples per loop of a kernel, from `ncu -i X.ncu-rep --page source --csv`.
usage: ncu -i rep --page source --csv > src.csv; tools/ncu_regions.py src.csv"""
import csv, re, sys

rows = list(csv.reader(open(sys.argv[1])))
hdr, data = rows[1], rows[2:]
ia, ie, it, isrc, ismp = (hdr.index(k) for k in ("Address", "Instructions Executed", "Thread Instructions Executed", "Source", "# Samples"))
base = int(data[0][ia], 16)
ins = [(int(r[ia], 16) - base, r[isrc].strip(), int(r[ie]), int(r[it]), int(r[ismp])) for r in data]
tot = sum(x[2] for x in ins); tots = sum(x[4] for x in ins)
loops = []
for a, op, e, t, s in ins:
    m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)*(0x[0-9a-f]+)", op)
    if m:
        tgt = int(m.group(1), 16) - base
        if 0 <= tgt <= a and a - tgt < 0x3000:
            loops.append((tgt, a))
covered = set()
print(f"total warp instructions {tot:.4g}, samples {tots}")
for a, b in loops:
    body = [x for x in ins if a <= x[0] <= b]
    e = sum(x[2] for x in body); s = sum(x[4] for x in body); t = sum(x[3] for x in body)
    if e / tot > 0.002:
        print(f"loop {a:#06x}..{b:#06x} {len(body):4d} instr: share {e/tot:6.3f} samples {s/tots:6.3f} lanes {t/max(e,1):5.1f} trips {e/len(body):.3g}")
    covered |= {x[0] for x in body}
seg, segs = None, []
for x in ins:
    if x[0] in covered:
        if seg: segs.append(seg); seg = None
        continue
    if seg is None: seg = [x[0], x[0], 0, 0, 0, 0]
    seg[1] = x[0]; seg[2] += x[2]; seg[3] += x[3]; seg[4] += x[4]; seg[5] += 1
if seg: segs.append(seg)
rest = sum(sg[2] for sg in segs)
print(f"outside loops: share {rest/tot:.3f}")
for sg in segs:
    if sg[2] / tot > 0.004:
        print(f"  seg {sg[0]:#06x}..{sg[1]:#06x} {sg[5]:4d} instr: share {sg[2]/tot:6.3f} samples {sg[4]/tots:6.3f} lanes {sg[3]/max(sg[2],1):5.1f}")
