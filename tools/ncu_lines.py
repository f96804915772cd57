"""Per-source-line instruction / stall-sample shares from an .ncu-rep (cuda,sass view)."""
import csv, subprocess, sys, io, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur = None; items = []
hdr = None
for r in csv.reader(io.StringIO(out)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Function Name": continue
    if r[0] == "Line No": hdr = r; ie = hdr.index("Instructions Executed"); ws = hdr.index("Warp Stall Sampling (All Samples)"); te = hdr.index("Avg. Threads Executed"); continue
    if hdr is None or r[2] != '-': continue     # only source-line rows (Address == '-')
    try: items.append((cur, int(r[0]), r[1].strip()[:95], int(r[ie] or 0), int(r[ws] or 0), r[te]))
    except Exception: pass
ti = sum(x[3] for x in items) or 1; ts = sum(x[4] for x in items) or 1
print(f"total inst {ti}  samples {ts}")
byfile = collections.Counter()
for x in items: byfile[x[0]] += x[3]
print({k: f"{100*v/ti:.1f}%" for k, v in byfile.items()})
for x in sorted(items, key=lambda t: -t[3])[:top]:
    print(f"{100*x[3]/ti:5.1f}% inst {100*x[4]/ts:5.1f}% smp thr {x[5]:>4s} {x[0]}:{x[1]:<4d} {x[2]}")
