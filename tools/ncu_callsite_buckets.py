"""Dynamic instructions of one kernel of an .ncu-rep bucketed by CALL SITE: every SASS instruction is attributed to the
line of a chosen source file that its inline chain passes through (helpers such as rget / rfind / ballots are charged to
the statement that called them, which the per-line source page cannot do).

  python tools/ncu_callsite_buckets.py <rep> <kernel-regex> <mangled-substring> <lib.so> <file.cuh> <units> [ranges.txt]

ranges.txt: lines "first last label" over <file.cuh>; default = one bucket per source line, top 40.
The library must be the build the capture ran (instruction count is checked)."""
import collections, csv, io, os, re, subprocess, sys, tempfile

rep, kre, mangled, lib, srcfile, units = sys.argv[1], sys.argv[2], sys.argv[3], sys.argv[4], sys.argv[5], float(sys.argv[6])
ranges = []
if len(sys.argv) > 7:
    for l in open(sys.argv[7]):
        l = l.split("#")[0].strip()
        if l:
            a, b, lab = l.split(None, 2)
            ranges.append((int(a), int(b), lab))

with tempfile.TemporaryDirectory() as tmp:
    subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(lib)], cwd=tmp, check=True, capture_output=True)
    cubin = [f for f in os.listdir(tmp) if f.endswith(".cubin")][0]
    sass = subprocess.run(["nvdisasm", "-gi", cubin], cwd=tmp, check=True, capture_output=True, text=True).stdout.split("\n")
start = end = None
for i, l in enumerate(sass):
    if l.startswith("//--------------------- .text.") and mangled in l and start is None:
        start = i
    elif start is not None and end is None and l.startswith("//--------------------- ") and i > start:
        end = i
static = []   # per instruction: (callsite line in srcfile or None, innermost (file, line))
chain = []
fresh = True
for l in sass[start:end]:
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        if fresh:
            chain = []
            fresh = False
        chain.append((os.path.basename(m.group(1)), int(m.group(2))))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,6}\*/\s+\S", l):
        fresh = True
        site = next((ln for f, ln in reversed(chain) if f == srcfile), None)   # outermost frame in the file
        static.append((site, chain[0] if chain else ("?", 0), chain[-1] if chain else ("?", 0)))

out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass", "--kernel-name", "regex:" + kre],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(out)))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
ie, te = hdr.index("Instructions Executed"), hdr.index("Thread Instructions Executed")
dyn = [(int(r[ie] or 0), int(r[te] or 0), r[1].strip()) for r in rows[hi + 1:] if len(r) > te]
if len(dyn) != len(static):
    print(f"WARNING: capture has {len(dyn)} SASS instructions, the library {len(static)}: not the same build", file=sys.stderr)
n = min(len(dyn), len(static))
tot = sum(d[0] for d in dyn[:n]) or 1
by = collections.defaultdict(lambda: [0, 0])
other = collections.defaultdict(lambda: [0, 0])
for (site, inner, outer), (ic, tc, _) in zip(static[:n], dyn[:n]):
    if site is None:
        other[outer][0] += ic; other[outer][1] += tc
    else:
        by[site][0] += ic; by[site][1] += tc
print(f"total warp instructions {tot}  per unit {tot / units:.1f}")
if ranges:
    for a, b, lab in ranges:
        s = sum(v[0] for k, v in by.items() if a <= k <= b)
        t = sum(v[1] for k, v in by.items() if a <= k <= b)
        print(f"{100 * s / tot:5.1f}%  {s / units:8.1f}/unit  lanes {t / max(s, 1):5.1f}  {srcfile}:{a}-{b}  {lab}")
    covered = sum(v[0] for k, v in by.items() if any(a <= k <= b for a, b, _ in ranges))
    print(f"{100 * (sum(v[0] for v in by.values()) - covered) / tot:5.1f}%  {srcfile} lines outside the ranges")
else:
    for k, v in sorted(by.items(), key=lambda kv: -kv[1][0])[:40]:
        print(f"{100 * v[0] / tot:5.1f}%  {v[0] / units:8.1f}/unit  lanes {v[1] / max(v[0], 1):5.1f}  {srcfile}:{k}")
print("not through", srcfile, "(by outermost line):")
for k, v in sorted(other.items(), key=lambda kv: -kv[1][0])[:25]:
    print(f"{100 * v[0] / tot:5.1f}%  {v[0] / units:8.1f}/unit  lanes {v[1] / max(v[0], 1):5.1f}  {k[0]}:{k[1]}")
