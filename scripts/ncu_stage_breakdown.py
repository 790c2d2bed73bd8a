"""Dynamic instruction breakdown of the fused kernel per pipeline stage / source file.
usage: ncu_stage_breakdown.py <ncu source-page sass csv> <nvdisasm -g -c output> <kernel substr> <warp-iterations>"""
import csv, re, collections, sys, os
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sass_csv, disasm, kern, iters = sys.argv[1], sys.argv[2], sys.argv[3], float(sys.argv[4])
addr2line={}; cur=None; infn=False
for ln in open(disasm, errors='replace'):
    m=re.match(r"\s*\.section\s+\.text\.(\S+)", ln)
    if m: infn = kern in m.group(1); cur=None; continue
    if not infn: continue
    m=re.search(r'//## File "([^"]+)", line (\d+)', ln)
    if m: cur=(m.group(1).split("/")[-1], int(m.group(2))); continue
    m=re.search(r"/\*([0-9a-f]{4,})\*/\s+(\S.*?);", ln)
    if m and cur: addr2line[int(m.group(1),16)]=(cur,m.group(2))
rows=list(csv.reader(open(sass_csv)))
hdr=next(i for i,r in enumerate(rows) if r and r[0]=="Address"); H=rows[hdr]
ia,ie=H.index("Address"),H.index("Instructions Executed")
base=None; per=collections.Counter(); tot=0; perop=collections.Counter()
for r in rows[hdr+1:]:
    if len(r)<=ie: continue
    try: a=int(r[ia],16)
    except ValueError: continue
    if base is None: base=a
    n=int(float(r[ie] or 0)); tot+=n
    key,op=addr2line.get(a-base,(("?",0),"?"))
    per[key]+=n
    op=re.sub(r"^@!?U?P\d+\s+","",op).split()[0].split(".")[0] if op!="?" else "?"
    perop[op]+=n
print("total",tot,"per warp-iteration %.1f" % (tot/iters))
src={}
for f in ['hydro_gen_b200/csrc/hg_fused_body.cuh','hydro_gen_b200/csrc/hg_cell.cuh','include/hg_defined_math.h','hydro_gen_b200/csrc/hg_fused.cu']:
    src[f.split('/')[-1]]=open(os.path.join(ROOT,f)).read().split('\n')
body=src['hg_fused_body.cuh']
marks=[(i+1,l.strip()) for i,l in enumerate(body) if '// ------------------------------------------------------------' in l]
cell=src['hg_cell.cuh']
cmarks=[(i+1,re.sub(r'\(.*','',l.split('hg_')[1])) for i,l in enumerate(cell) if l.startswith('HG_FN') and 'hg_' in l]
def stage(f,l):
    if f=='hg_fused_body.cuh':
        name='body:pre'
        for ln,t in marks:
            if l>=ln: name='body:'+t.replace('-','').replace('/','').strip()
        return name
    if f=='hg_cell.cuh':
        name='cell:?'
        for ln,t in cmarks:
            if l>=ln: name='cell:'+t
        return name
    return f
agg=collections.Counter()
for (f,l),n in per.items(): agg[stage(f,l)]+=n
for k,v in agg.most_common(): print(f"  {k:40s} {v/iters:7.1f} /iter  {100*v/tot:5.1f}%")
print("by opcode:", "  ".join(f"{k} {v/iters:.0f}" for k,v in perop.most_common(34)))
