"""Summarise an .ncu-rep: key raw metrics and the hottest source lines (needs -lineinfo)."""
import csv, subprocess, sys, collections, io
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
want = ['Kernel Name', 'gpu__time_duration.sum', 'smsp__inst_executed.sum', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__occupancy_limit_registers',
        'launch__occupancy_limit_shared_mem', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum', 'smsp__inst_executed_op_shared_ld.sum', 'smsp__inst_executed_op_shared_st.sum',
        'smsp__inst_executed_op_local_ld.sum', 'smsp__inst_executed_op_local_st.sum',
        'sm__inst_executed_pipe_fma.sum', 'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_fp64.sum', 'sm__inst_executed_pipe_lsu.sum',
        'sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__cycles_active.avg']
want += [h for h in hdr if 'issue_stalled' in h and 'per_issue_active' in h]
for r in rows[2:]:
    for w in want:
        if w in hdr:
            i = hdr.index(w)
            print(f'{w:92s} {r[i]:>16s} {units[i]}')
    print()
if len(sys.argv) > 2:
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    # find header row
    hi = next(i for i, r in enumerate(rows) if 'Source' in r and any('Instructions Executed' in c for c in r))
    hdr = rows[hi]
    si = hdr.index('Source'); 
    ie = next(i for i, c in enumerate(hdr) if c.strip() == 'Instructions Executed')
    ws = next((i for i, c in enumerate(hdr) if c.strip().startswith('Warp Stall Sampling (All')), None)
    li = hdr.index('#') if '#' in hdr else 0
    items = []
    for r in rows[hi + 1:]:
        try:
            items.append((int(r[ie] or 0), int(r[ws] or 0) if ws is not None else 0, r[li], r[si][:110]))
        except Exception:
            pass
    tot_i = sum(x[0] for x in items) or 1; tot_s = sum(x[1] for x in items) or 1
    print("total inst", tot_i, "total samples", tot_s)
    print("--- top lines by stall samples")
    for x in sorted(items, key=lambda t: -t[1])[:int(sys.argv[2])]:
        print(f"{100*x[1]/tot_s:5.1f}% smp {100*x[0]/tot_i:5.1f}% inst  L{x[2]:>5s} {x[3]}")
