"""Static SASS instruction count of one kernel per INNERMOST source line (inlined helpers attributed to their own lines),
split by function body / out-of-line callee.  python tools/sass_innermost.py <kernel-substring> [min-count]"""
import re, collections, subprocess, sys, os, tempfile
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
kern = sys.argv[1]; mn = int(sys.argv[2]) if len(sys.argv) > 2 else 10
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "surtr_b200", "libsurtr_b200.so")], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-gi", cubin], cwd=tmp, check=True, capture_output=True, text=True).stdout.split("\n")
start = end = None
for i, l in enumerate(sass):
    if l.startswith("//--------------------- .text.") and kern in l: start = i
    elif start is not None and end is None and l.startswith("//--------------------- ") and i > start: end = i
sec = sass[start:end]
cur = None; c = collections.Counter(); func = 'body'; fc = collections.Counter(); chain = False; ops = collections.Counter()
for l in sec:
    m = re.search(r'//## File "([^"]+)", line (\d+)(.*)', l)
    if m:
        if 'inlined at' in m.group(3):
            if not chain: cur = (m.group(1).split('/')[-1], int(m.group(2)))
            chain = True
        else:
            if chain: chain = False
            else: cur = (m.group(1).split('/')[-1], int(m.group(2)))
        continue
    m2 = re.match(r"\s*\$\S*\$(_ZN\w+|__internal\w+)\S*:", l)
    if m2: func = m2.group(1)[:48]
    m3 = re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", l)
    if m3:
        chain = False
        c[(func,) + (cur or ('?', 0))] += 1
        fc[func] += 1
        ops[m3.group(1).split('.')[0]] += 1
print(dict(fc), "total", sum(fc.values()), "=", sum(fc.values()) * 16 // 1024, "KB")
print("opcodes:", ", ".join(f"{k} {v}" for k, v in ops.most_common(25)))
srcs = {}
for (fn, f, ln), v in sorted(c.items(), key=lambda t: (t[0][0] != 'body', t[0][1], t[0][2])):
    if v < mn: continue
    if f not in srcs:
        pth = os.path.join(ROOT, "surtr_b200", "csrc", f)
        srcs[f] = open(pth).read().split("\n") if os.path.exists(pth) else []
    txt = srcs[f][ln - 1].strip()[:80] if 0 < ln <= len(srcs[f]) else ""
    print(f"{v:5d} {fn[:14]:14s} {f}:{ln:<4d} {txt}")
