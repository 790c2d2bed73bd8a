"""Attribute ncu per-SASS-instruction counts to source lines.
usage: ncu_lines.py <sass_page.csv> <nvdisasm -g -c output> [kernel substring]"""
import csv, re, sys, collections
sass_csv, disasm = sys.argv[1], sys.argv[2]
kern = sys.argv[3] if len(sys.argv) > 3 else "k_fused_step"
# 1. address -> line from nvdisasm (only inside the wanted function)
addr2line = {}
cur = None; infn = False
for ln in open(disasm, errors="replace"):
    m = re.match(r"\s*\.section\s+\.text\.(\S+)", ln)
    if m:
        infn = kern in m.group(1); cur = None; continue
    if not infn: continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2))); continue
    m = re.search(r"/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m and cur:
        addr2line[int(m.group(1), 16)] = (cur, m.group(2))
rows = list(csv.reader(open(sass_csv)))
hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
H = rows[hdr]
ia, ie, isamp = H.index("Address"), H.index("Instructions Executed"), H.index("# Samples")
base = None
per_line = collections.Counter(); samp_line = collections.Counter(); per_op = collections.Counter()
tot = 0; tots = 0
for r in rows[hdr + 1:]:
    if len(r) <= ie: continue
    try: a = int(r[ia], 16) if r[ia].startswith("0x") or re.fullmatch(r"[0-9a-fA-F]+", r[ia]) else int(r[ia])
    except ValueError: continue
    if base is None: base = a
    off = a - base
    n = int(float(r[ie] or 0)); s = int(float(r[isamp] or 0))
    tot += n; tots += s
    key, op = addr2line.get(off, (("?", 0), "?"))
    per_line[key] += n; samp_line[key] += s
    per_op[op.split()[0].split(".")[0] if op != "?" else "?"] += n
print(f"total warp-instructions {tot}, samples {tots}")
print("top lines by executed instructions (share, stall-sample share):")
for k, v in per_line.most_common(45):
    print(f"  {k[0]}:{k[1]:<5d} {100*v/tot:5.1f}%   samples {100*samp_line[k]/max(tots,1):5.1f}%")
print("by opcode:")
for k, v in per_op.most_common(25):
    print(f"  {k:10s} {100*v/tot:5.1f}%")
# per-stage aggregation (line ranges of the current sources)
ranges = [("cell:flux", "hg_cell.cuh", 60, 100), ("cell:erosion", "hg_cell.cuh", 101, 158), ("cell:back/bilerp", "hg_cell.cuh", 159, 181),
          ("cell:thermal_outflow", "hg_cell.cuh", 182, 235), ("cell:thermal_delta", "hg_cell.cuh", 236, 245), ("cell:smooth", "hg_cell.cuh", 246, 274),
          ("fused:prologue", "hg_fused.cu", 146, 197), ("fused:L", "hg_fused.cu", 198, 218), ("fused:A", "hg_fused.cu", 219, 246),
          ("fused:B", "hg_fused.cu", 247, 270), ("fused:CD", "hg_fused.cu", 271, 319), ("fused:EF", "hg_fused.cu", 320, 368),
          ("fused:G+loop", "hg_fused.cu", 369, 400), ("defined_math", "hg_defined_math.h", 0, 999)]
agg = collections.Counter(); sagg = collections.Counter()
for (f, l), v in per_line.items():
    name = next((n for n, ff, a, b in ranges if ff == f and a <= l <= b), f"other:{f}")
    agg[name] += v; sagg[name] += samp_line[(f, l)]
print("by stage:")
for k, v in agg.most_common():
    print(f"  {k:24s} {100*v/tot:5.1f}%  ({v/668160:6.1f} instr/warp-row)   samples {100*sagg[k]/max(tots,1):5.1f}%")
