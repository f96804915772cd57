#!/bin/bash
# Turn the files of one tools/r2_evidence.sh run (gpurun_out/<tag>_*) into the committed round-2 evidence under profiles/.
# The library in the tree must be the build the run used (instruction buckets correlate the capture with its SASS).
T=${1:?tag}
G=gpurun_out
P=profiles
set -e
python tools/rowg_table.py $T > $P/r2_kernel_rooflines.md
grep -v "^==" $G/${T}_launches.csv > $P/r2_launches.csv
python tools/launch_table.py $P/r2_launches.csv > $P/r2_launch_table.txt
tail -1 $G/${T}_bench.json > $P/r2_bench_line.json
tail -1 $G/${T}_bench_ref.json > $P/r2_bench_reference_arm.json
N4=$(python -c "import json;print([json.loads(l) for l in open('$G/${T}_k3_cfg4.log') if l.startswith('{')][-1]['candidates'])")
F4=$(python -c "import json;print([json.loads(l) for l in open('$G/${T}_k4_cfg4.log') if l.startswith('{')][-1]['fragments'])")
python tools/ncu_k.py $G/${T}_k3_cfg4.ncu-rep clip_fast_kernel $N4 40 > $P/r2_k3_metrics.txt
python tools/ncu_k.py $G/${T}_k4_cfg4.ncu-rep assemble_gather $F4 40 > $P/r2_k4_gather_metrics.txt
python tools/ncu_k.py $G/${T}_k3_cfg2.ncu-rep clip_fast_kernel 4096 25 > $P/r2_k3_config2_metrics.txt
cat > /tmp/k3_ranges.txt <<'R'
312 364 prefilter (box vs planes) + setup
365 382 plane queue: peek / pop
383 409 classify + no-cut exits
410 428 straddle loop (ring slots of clipped lanes)
429 447 prefix by ballots
448 462 overflow / compaction trigger
463 472 list write
473 491 insert new vertices
492 522 patch walk
523 526 probe check
527 537 compose rings
538 548 seq dispatch
549 575 live update, refresh
110 242 sequential replay (fast_seq_cut)
243 284 compaction (fast_compact)
285 310 all-in-plane box test
R
python tools/ncu_callsite_buckets.py $G/${T}_k3_cfg4.ncu-rep clip_fast_kernel clip_fast_kernelILi2ELb0ELi4ELb1 surtr_b200/libsurtr_b200.so clip_fast.cuh $N4 /tmp/k3_ranges.txt > $P/r2_k3_instruction_buckets.txt
python tools/ncu_callsite_buckets.py $G/${T}_k3_cfg4.ncu-rep clip_fast_kernel clip_fast_kernelILi2ELb0ELi4ELb1 surtr_b200/libsurtr_b200.so kernels.cuh $N4 2>/dev/null | head -24 >> $P/r2_k3_instruction_buckets.txt || true
python tools/ncu_traffic.py $G/${T}_k3_cfg4.ncu-rep $G/${T}_k3_cfg4.log clip_fast_kernel
python tools/sass_histogram.py > $P/r2_sass_opcodes.txt
/usr/local/cuda/bin/nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -fmad=false -Xcompiler -fPIC -shared -Xptxas -v -o /tmp/ptxas_check.so surtr_b200/csrc/surtr_engine.cu 2>&1 | grep -E "Compiling entry|Used|spill" | sed 's/ptxas info    : //' > $P/r2_ptxas_resources.txt
echo ok
cp $G/configs_r2.json $P/r2_configs.json 2>/dev/null || true
grep -v "^\[surtr\]       " $G/${T}_dofracture_trace.txt | tail -40 > $P/r2_dofracture_trace.txt
for w in config4 config3 config2 mesh; do tail -1 $G/${T}_phases_$w.jsonl; done > $P/r2_phase_times.jsonl
