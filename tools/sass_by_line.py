"""Static code size per source line of one kernel: extract the cubin from the built library, disassemble with line
info and count SASS instructions per innermost source line (instruction-cache footprint; see profiles/README.md).

  python tools/sass_by_line.py [kernel-substring] [n-top]
"""
import collections, os, re, subprocess, sys, tempfile

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
kern = sys.argv[1] if len(sys.argv) > 1 else "clip_sub_kernelILi32"
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 30
with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.join(ROOT, "surtr_b200", "libsurtr_b200.so")], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-gi", cubin], cwd=tmp, check=True, capture_output=True, text=True).stdout.split("\n")
start = end = None
for i, l in enumerate(sass):
    if l.startswith("//--------------------- .text.") and kern in l:
        start = i
    elif start is not None and end is None and l.startswith("//--------------------- ") and i > start:
        end = i
sec = sass[start:end]
cnt, fcnt = collections.Counter(), collections.Counter()
cur, pending_inline, func = None, False, "kernel body"
for l in sec:
    m = re.search(r'//## File "([^"]+)", line (\d+)( inlined at)?', l)
    if m:
        if m.group(3):
            cur, pending_inline = (os.path.basename(m.group(1)), int(m.group(2))), True
        elif pending_inline:
            pending_inline = False          # the call-site line that follows an "inlined at" marker
        else:
            cur = (os.path.basename(m.group(1)), int(m.group(2)))
        continue
    m2 = re.match(r"\s*\$\S*\$(_ZN\w+|__internal\w+)\S*:", l)
    if m2:
        func = m2.group(1)[:60]
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+\S", l):
        pending_inline = False
        cnt[cur] += 1
        fcnt[func] += 1
print("kernel", kern, "SASS instructions", sum(cnt.values()), "=", sum(cnt.values()) * 16 // 1024, "KB")
for f, c in fcnt.most_common():
    print(f"  {c:6d}  {f}")
files = collections.Counter()
for (f, ln), c in cnt.items():
    files[f] += c
print("by file:", dict(files.most_common()))
print("top lines:")
for (f, ln), c in sorted(cnt.items(), key=lambda kv: -kv[1])[:ntop]:
    print(f"  {c:5d}  {f}:{ln}")
