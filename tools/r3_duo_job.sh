# A/B of the two-pairs-per-warp small tier (SURTR_K3_DUO=1) against the shipped one-warp-per-pair kernel: parity suite, then phase times
O=gpurun_out; T=${1:-r3b}
SURTR_K3_DUO=1 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > $O/${T}_pytest_duo.log; cat $O/${T}_pytest_duo.log
for rep in 1 2; do for w in config4 config3 config2; do
  EVENTS=256 python tools/gpu_profile_workloads.py $w 1 2>&1 | tail -1 >> $O/${T}_phases_${w}_fast.jsonl
  SURTR_K3_DUO=1 EVENTS=256 python tools/gpu_profile_workloads.py $w 1 2>&1 | tail -1 >> $O/${T}_phases_${w}_duo.jsonl
done; done
python - <<P
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_phases_*.jsonl')):
    for l in open(f):
        try: d=json.loads(l)
        except Exception: print(f, l[:200]); continue
        k=d['kernel_ms_unprofiled']
        print(f.split('phases_')[1], d['fragments'], d['seq_cuts'], d['tier1b'], 'k3', round(k['k3_clip_small'],4), 'k3large', round(k['k3_clip_large'],4), 'k4', round(k['k4_gather'],4))
P
