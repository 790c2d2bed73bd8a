#!/usr/bin/env python
"""Static SASS statistics per loop of a kernel: every backward branch span (outermost first), instruction count and
opcode histogram.  usage: sass_loops.py <obj|cubin> <kernel substring> [min_instructions]"""
import collections, re, subprocess, sys
obj, kern = sys.argv[1], sys.argv[2]
minlen = int(sys.argv[3]) if len(sys.argv) > 3 else 150
out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
ins = []; on = False
for ln in out.splitlines():
    if "Function :" in ln:
        on = kern in ln
        continue
    if not on: continue
    m = re.match(r"\s+/\*([0-9a-f]{4,})\*/\s+(.*?);", ln)
    if m: ins.append((int(m.group(1), 16), m.group(2).strip()))
print(f"{kern}: {len(ins)} SASS instructions")
def opname(t):
    t = re.sub(r"^@!?U?P\d+\s+", "", t)
    return t.split()[0].split(".")[0]
loops = []
for a, t in ins:
    m = re.search(r"\bBRA(?:\.U)?\b.*?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a: loops.append((tgt, a))
loops.sort(key=lambda l: (l[0], -l[1]))
for lo, hi in loops:
    sel = [x for x in ins if lo <= x[0] <= hi]
    if len(sel) < minlen: continue
    h = collections.Counter(opname(t) for _, t in sel)
    calls = sum(1 for _, t in sel if "CALL" in t)
    print(f"loop [{lo:#x},{hi:#x}]: {len(sel)} instructions ({len(sel) * 16 / 1024:.1f} KB)")
    print("   " + "  ".join(f"{k} {v}" for k, v in h.most_common(40)))
