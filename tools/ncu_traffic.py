"""profiles/r2_k3_traffic.json from an `ncu --set full` capture of the K3 small-tier launch (bench.py's roofline.traffic):
  python tools/ncu_traffic.py <rep> <workload-json-line-file> [kernel-regex]
The workload line is what tools/gpu_profile_workloads.py printed for the captured run (candidates, algorithmic bytes).
Records the sha of the kernel sources the capture was built from: bench.py uses the figure only while that still matches."""
import csv, io, json, os, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench_secondary
rep, meta_file = sys.argv[1], sys.argv[2]
kre = sys.argv[3] if len(sys.argv) > 3 else "clip_fast_kernel"
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--kernel-name", "regex:" + kre], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
sc = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
def val(n):
    i = hdr.index(n)
    return float(r[i].replace(",", "")) * sc.get(units[i], 1)
meta = [json.loads(l) for l in open(meta_file) if l.startswith("{")][-1]
rd, wr = val("dram__bytes_read.sum"), val("dram__bytes_write.sum")
out = {"kernel": r[hdr.index("Kernel Name")], "workload": f"{meta['workload']}: {meta['pairs']} pairs, {meta['candidates']} candidates, {meta['fragments']} fragments per launch",
       "source": f"ncu --set full --clock-control none ({os.path.basename(rep)}), caches flushed before the launch: cold-cache traffic",
       "launches": 1, "dram_bytes_read": int(rd), "dram_bytes_write": int(wr), "dram_bytes_per_launch": int(rd + wr),
       "algorithmic_bytes_per_launch": int(meta["k3_algorithmic_bytes"]), "candidates_per_launch": int(meta["candidates"]),
       "kernel_source_sha": bench_secondary.kernel_source_sha()}
json.dump(out, open(os.path.join(ROOT, "profiles", "r2_k3_traffic.json"), "w"), indent=1)
print(json.dumps(out))
