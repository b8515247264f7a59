"""Compact table of registers / spills per kernel from the *.ptxas.log files of the last build.  usage: ptxas_table.py [dir]"""
import glob, os, re, subprocess, sys
d = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(__file__), "..", "globalredundancyresolution_b200", "csrc")
for f in sorted(glob.glob(os.path.join(d, "*.ptxas.log"))):
    if ".dbg." in f:
        continue
    cur, spill = None, None
    for line in open(f):
        m = re.search(r"Compiling entry function '(\S+)'", line)
        if m:
            cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", cur).replace("void grr::", "")
        m = re.search(r"(\d+) bytes spill stores", line)
        if m:
            spill = m.group(1)
        m = re.search(r"Used (\d+) registers", line)
        if m and cur:
            print("%-22s %-48s regs %3s spill %s" % (os.path.basename(f).replace(".ptxas.log", ""), cur, m.group(1), spill))
            cur = None
