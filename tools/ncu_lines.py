"""Stall samples / executed instructions per CUDA source line from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass > f.csv`
(needs -lineinfo and --import-source on).  usage: tools/ncu_lines.py f.csv [top]"""
import csv, sys
rows = list(csv.reader(open(sys.argv[1])))
top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
cur, out, tot_s, tot_i = None, [], 0, 0
for r in rows:
    if len(r) >= 2 and r[0] == "File Path":
        cur = r[1].split("/")[-1]; continue
    if len(r) < 9 or r[0] in ("Line No", "Function Name", ""):
        continue
    try:
        ln, smp, ins, thr = int(r[0]), int(r[6]), int(r[7]), int(r[8])
    except ValueError:
        continue
    out.append((smp, ins, thr, cur, ln, r[1].strip()[:110])); tot_s += smp; tot_i += ins
print("total samples", tot_s, "warp instructions", tot_i)
for smp, ins, thr, f, ln, src in sorted(out, reverse=True)[:top]:
    print("%5.2f%% smp %5.2f%% ins lanes %4.1f  %s:%d  %s" % (100 * smp / tot_s, 100 * ins / tot_i, thr / max(ins, 1), f, ln, src))
