#!/bin/bash
# Round-2 evidence run on one B200 (gpurun): parity suite, per-kernel roofline table (row g), bench lines of both arms,
# ncu launch list of the bench command, full captures of K3 and K4.  Usage: tools/r2_evidence.sh <tag>
T=${1:-r2z}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1
tail -1 $O/${T}_pytest.log
# ---- per-kernel table: every kernel of one event through the blob path, cold caches (ncu flushes between replays)
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,lts__t_bytes.sum,sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active,sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active,smsp__inst_executed.sum,smsp__thread_inst_executed_per_inst_executed.ratio,sm__warps_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__registers_per_thread,launch__grid_size,launch__block_size
for w in config4 config3; do
  BLOB=1 EVENTS=256 ncu --profile-from-start off --clock-control none --metrics $M --csv --log-file $O/${T}_rowg_${w}.csv \
    python tools/gpu_profile_workloads.py $w 1 > $O/${T}_rowg_${w}.json 2>&1
done
# ---- unprofiled phase times of the same workloads (CUDA events)
for rep in 1 2 3; do for w in config4 config3 config2 mesh; do EVENTS=256 python tools/gpu_profile_workloads.py $w 1 2>&1 | tail -1 >> $O/${T}_phases_$w.jsonl; done; done
# ---- full capture of the dominant kernel on one resident batch of the bench (512 events) FIRST: its cold DRAM traffic is what
#      the bench line reports as roofline.traffic (profiles/r2_k3_traffic.json, tied to the kernel sources by their hash)
EVENTS=512 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:clip_fast_kernel -c 1 -f -o $O/${T}_k3_cfg4 \
  python tools/gpu_profile_workloads.py config4 1 > $O/${T}_k3_cfg4.log 2>&1
python tools/ncu_traffic.py $O/${T}_k3_cfg4.ncu-rep $O/${T}_k3_cfg4.log clip_fast_kernel > $O/${T}_k3_traffic.log 2>&1
# ---- bench lines (never under a profiler)
python bench.py > $O/${T}_bench.json 2> $O/${T}_bench.err
python bench.py --impl reference --steps 3 --warmup 1 > $O/${T}_bench_ref.json 2> $O/${T}_bench_ref.err
# ---- launch list of the bench command
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/${T}_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-secondary --no-cpu-baseline --events 1024 > $O/${T}_ncu_bench.log 2>&1
# ---- full captures of K4 on the same batch, and of K3 on config 2
EVENTS=512 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:assemble_gather -c 1 -f -o $O/${T}_k4_cfg4 \
  python tools/gpu_profile_workloads.py config4 1 > $O/${T}_k4_cfg4.log 2>&1
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:clip_fast_kernel -c 1 -f -o $O/${T}_k3_cfg2 \
  python tools/gpu_profile_workloads.py config2 1 > $O/${T}_k3_cfg2.log 2>&1
python tests/measure/gpu_dofracture_trace.py > /dev/null 2> $O/${T}_dofracture_trace.txt
python tests/measure/gpu_configs.py 256 256 > $O/${T}_configs.log 2>&1
echo done
