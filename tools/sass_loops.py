#!/usr/bin/env python
"""List the loops (backward branches) of one kernel in the built library with their length in SASS
instructions -- the static figure quoted in profiles/*_edge_kernel.md next to ncu's executed counts.
usage: tools/sass_loops.py <kernel regex> [lib]"""
import re, subprocess, sys

def main():
    pat = re.compile(sys.argv[1])
    lib = sys.argv[2] if len(sys.argv) > 2 else "globalredundancyresolution_b200/libgrr_b200.so"
    txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    cur, funcs = None, {}
    for line in txt.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = m.group(1); funcs[cur] = []; continue
        m = re.match(r"\s*/\*([0-9a-f]{4,})\*/\s+(.*?);", line)
        if m and cur:
            funcs[cur].append((int(m.group(1), 16), m.group(2).strip()))
    for name, ins in funcs.items():
        if not pat.search(name):
            continue
        print(f"{name}: {len(ins)} instructions")
        for addr, op in ins:
            m = re.search(r"BRA(?:\.U)?\s+(?:!?U?P\d+,\s*)*(0x[0-9a-f]+)", op)
            if m and int(m.group(1), 16) <= addr:
                tgt = int(m.group(1), 16)
                body = [o for a, o in ins if tgt <= a <= addr]
                kinds = {}
                for o in body:
                    k = re.sub(r"^@!?U?P\d+\s+", "", o).split()[0].split(".")[0]
                    kinds[k] = kinds.get(k, 0) + 1
                top = ", ".join(f"{k} {v}" for k, v in sorted(kinds.items(), key=lambda kv: -kv[1])[:12])
                print(f"  loop {tgt:#06x}..{addr:#06x}: {len(body):4d} instr  [{top}]")

if __name__ == "__main__":
    main()
