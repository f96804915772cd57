#!/bin/bash
# A/B of the small tier's launch shape: SURTR_K3_WARPS = 2 (two pairs per block), 1 (one), 0 (persistent warps + ticket).
T=${1:-r2n}
O=gpurun_out
mkdir -p $O
python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1
tail -1 $O/${T}_pytest.log
for rep in 1 2 3; do
for k in 2 0; do
  for w in config4 config3 config2; do
    SURTR_K3_WARPS=$k EVENTS=256 python tools/gpu_profile_workloads.py $w 1 2>&1 | tail -1 >> $O/${T}_k3warps${k}_$w.jsonl
  done
done
done
python tools/gpu_profile_workloads.py mesh 1 2>&1 | tail -1 > $O/${T}_mesh.json
python bench.py --no-cpu-baseline > $O/${T}_bench.json 2> $O/${T}_bench.err
echo done
