import csv, subprocess, sys, io, collections, re
rep = sys.argv[1]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur=None; hdr=None; rows=[]
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0]=="File Path": cur=r[1].split('/')[-1]; continue
    if r[0]=="Function Name": continue
    if r[0]=="Line No": hdr=r; ie=hdr.index("Instructions Executed"); te=hdr.index("Thread Instructions Executed"); continue
    if hdr is None or r[2] != '-': continue
    try: rows.append((cur,int(r[0]),int(r[ie] or 0),int(r[te] or 0), r[1].strip()[:60]))
    except: pass
tot=sum(x[2] for x in rows)
def show(name, pred):
    s=sum(x[2] for x in rows if pred(x)); t=sum(x[3] for x in rows if pred(x))
    print(f"{name:40s} {100*s/tot:5.1f}%  inst/pair {s/4096:8.0f}  lanes {t/max(s,1):5.1f}")
src=open('/root/repo/surtr_b200/csrc/clip_sub.cuh').read().split('\n')
def line_of(pat, start=0):
    for i,l in enumerate(src[start:], start+1):
        if pat in l: return i
    return None
marks=[("helpers", 1, line_of("struct CutState")),
       ("seq/compact/box fns", line_of("struct CutState"), line_of("__device__ int sub_clip_by_planes")),
       ("clip: setup+classify", line_of("__device__ int sub_clip_by_planes"), line_of("// ---- the plane cuts")),
       ("clip: straddle+scan+list", line_of("// ---- the plane cuts"), line_of("// insert: one new vertex per lane")),
       ("clip: insert", line_of("// insert: one new vertex per lane"), line_of("// patch (Poly.cpp:365-431)")),
       ("clip: patch walk+probe+compose", line_of("// patch (Poly.cpp:365-431)"), line_of("// lazy compaction: clipped vertices leave")),
       ("clip: live-mask update+refresh", line_of("// lazy compaction: clipped vertices leave"), line_of("// Poly::ExtractFaces + Poly::Moments in the reference")),
       ("moments", line_of("// Poly::ExtractFaces + Poly::Moments in the reference"), 100000)]
for name,a,b in marks:
    show("clip_sub.cuh "+name, lambda x,a=a,b=b: x[0]=='clip_sub.cuh' and a<=x[1]<b)
for f in sorted(set(x[0] for x in rows)):
    if f!='clip_sub.cuh': show(f, lambda x,f=f: x[0]==f)
print("total inst/pair", tot/4096)
