#!/usr/bin/env python
"""Static SASS statistics of a kernel in an object file: total instructions, opcode histogram,
and the same for the largest loop (backward branch span) = the steady-state body.
usage: sass_stats.py <obj> <kernel substring> [--div N]  (N = iterations per loop trip, e.g. 6)"""
import collections, re, subprocess, sys
obj, kern = sys.argv[1], sys.argv[2]
div = int(sys.argv[sys.argv.index("--div") + 1]) if "--div" in sys.argv else 1
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
# largest backward branch span
best = None
for a, t in ins:
    m = re.search(r"\bBRA(?:\.U)?\b.*?(0x[0-9a-f]+)", t)
    if m:
        tgt = int(m.group(1), 16)
        if tgt < a and (best is None or a - tgt > best[1] - best[0]): best = (tgt, a)
def hist(sel, title, d=1):
    h = collections.Counter(opname(t) for _, t in sel)
    n = len(sel)
    print(f"{title}: {n} instructions" + (f" = {n / d:.0f} per iteration" if d > 1 else ""))
    print("  " + "  ".join(f"{k} {v / d:.0f}" for k, v in h.most_common(28)))
if best:
    hist([x for x in ins if best[0] <= x[0] <= best[1]], f"largest loop [{best[0]:#x},{best[1]:#x}]", div)
hist(ins, "whole kernel")
