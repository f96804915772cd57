"""Key metrics + per-source-line dynamic instruction shares of ONE kernel of an .ncu-rep.
  python tools/ncu_k.py <rep> <kernel-regex> <units (e.g. candidates)> [top]"""
import csv, io, subprocess, sys, collections
rep, kern, units = sys.argv[1], sys.argv[2], float(sys.argv[3])
top = int(sys.argv[4]) if len(sys.argv) > 4 else 50
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr = rows[0]
want = ['Kernel Name', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'dram__bytes_read.sum',
        'dram__bytes_write.sum', 'launch__registers_per_thread', 'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum',
        'sm__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__grid_size']
want += [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h and 'not_issued' not in h]
r = rows[2]
for w in want:
    if w in hdr:
        print(f'{w:92s} {r[hdr.index(w)]:>18s} {rows[1][hdr.index(w)]}')
inst = float(r[hdr.index('smsp__inst_executed.sum')].replace(',', ''))
print(f"warp instructions per unit: {inst / units:.1f}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", "regex:" + kern], capture_output=True, text=True).stdout
cur = None; h2 = None; items = []
for r in csv.reader(io.StringIO(src)):
    if not r: continue
    if r[0] == "File Path": cur = r[1].split('/')[-1]; continue
    if r[0] == "Line No": h2 = r; ie = h2.index("Instructions Executed"); ws = h2.index("Warp Stall Sampling (All Samples)"); te = h2.index("Avg. Threads Executed"); continue
    if h2 is None or len(r) <= ie or r[2] != '-': continue
    try: items.append((cur, int(r[0]), r[1].strip()[:96], int(r[ie] or 0), int(r[ws] or 0), r[te]))
    except Exception: pass
ti = sum(x[3] for x in items) or 1; ts = sum(x[4] for x in items) or 1
byfile = collections.Counter()
for x in items: byfile[x[0]] += x[3]
print("source-page instructions", ti, {k: f"{100 * v / ti:.1f}%" for k, v in byfile.items()})
for x in sorted(items, key=lambda t: -t[3])[:top]:
    print(f"{100 * x[3] / ti:5.1f}% {x[3] / units:7.1f}/unit {100 * x[4] / ts:5.1f}% smp  {x[0]}:{x[1]:<4d} {x[2]}")
