"""Per-kernel roofline table (SURVEY section 8 row g) from the ncu CSVs tools/r2_evidence.sh wrote:
  python tools/rowg_table.py <tag> [peak GB/s]  ->  markdown on stdout
Reads gpurun_out/<tag>_rowg_<workload>_bulk0.{csv,json}: metrics per launch (ncu flushes the caches between replays, so
dram bytes are cold-cache), algorithmic bytes per kernel from the workload's own counts (tools/gpu_profile_workloads.py)."""
import csv, io, json, sys, os
tag = sys.argv[1]
peak = float(sys.argv[2]) if len(sys.argv) > 2 else json.load(open("MEASURED_PEAKS.json")).get("hbm_gbs", 6555.8)
sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
tsc = {"ns": 1e-3, "us": 1.0, "ms": 1e3, "s": 1e6}
for w in ("config4", "config3"):
    f = f"gpurun_out/{tag}_rowg_{w}_bulk0"
    if not os.path.exists(f + ".csv"):
        f = f"gpurun_out/{tag}_rowg_{w}"
    txt = open(f + ".csv").read()
    rows = list(csv.DictReader(io.StringIO(txt[txt.index('"ID"'):])))
    meta = [json.loads(l) for l in open(f + ".json") if l.startswith("{")][-1]
    alg = meta["algorithmic_bytes_per_kernel"]
    ker, order = {}, []
    for r in rows:
        key = (r["ID"], r["Kernel Name"])
        if key not in ker:
            ker[key] = {}
            order.append(key)
        ker[key][r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    print(f"\n### {w}: pairs {meta['pairs']}, candidates {meta['candidates']}, fragments {meta['fragments']} "
          f"(one event through surtr_upload_blob -> kernels -> surtr_download_blob_async)\n")
    print("| kernel | time us | share | algorithmic MB | alg GB/s | frac of %.1f GB/s | dram rd+wr MB (cold) | dram/alg | fma pipe %% | alu pipe %% | issue active %% | warps active %% | lanes/inst | regs | grid x block |" % peak)
    print("|---|---|---|---|---|---|---|---|---|---|---|---|---|---|---|")
    tot = sum(ker[k]["gpu__time_duration.sum"][0] * tsc[ker[k]["gpu__time_duration.sum"][1]] for k in order if "surtr::" in k[1])
    for key in order:
        if "surtr::" not in key[1]:
            continue
        m = ker[key]
        g = lambda n: m[n][0]
        dur = g("gpu__time_duration.sum") * tsc[m["gpu__time_duration.sum"][1]]
        rd = g("dram__bytes_read.sum") * sc[m["dram__bytes_read.sum"][1]]
        wr = g("dram__bytes_write.sum") * sc[m["dram__bytes_write.sum"][1]]
        name = key[1].split("(")[0].replace("void ", "").replace("surtr::", "")
        a = alg.get(name.split("<")[0])
        if name.startswith("clip_fast_kernel<4") or a is None:
            a = None
        print(f"| `{name}` | {dur:.1f} | {100 * dur / tot:.1f}% | " + (f"{a / 1e6:.2f} | {a / dur / 1e3:.0f} | {a / dur / 1e3 / peak:.4f}" if a else "- | - | -") +
              f" | {rd / 1e6:.2f}+{wr / 1e6:.2f} | " + (f"{(rd + wr) / a:.2f}" if a else "-") +
              f" | {g('sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active'):.1f} | {g('sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active'):.1f}"
              f" | {g('smsp__issue_active.avg.pct_of_peak_sustained_active'):.1f} | {g('sm__warps_active.avg.pct_of_peak_sustained_active'):.1f}"
              f" | {g('smsp__thread_inst_executed_per_inst_executed.ratio'):.1f} | {int(g('launch__registers_per_thread'))} | {int(g('launch__grid_size'))} x {int(g('launch__block_size'))} |")
    print(f"\nsum of kernel times {tot:.1f} us (serialised, cold caches under ncu; the unprofiled event: "
          f"{1e3 * sum(meta['kernel_ms_unprofiled'].values()):.1f} us)")
