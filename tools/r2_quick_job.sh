O=gpurun_out; T=${1:-r2q}
python -m pytest tests -m gpu -x -q 2>&1 | tail -1 > $O/${T}_pytest.log; cat $O/${T}_pytest.log
for rep in 1 2; do for w in config4 config3 config2; do EVENTS=256 python tools/gpu_profile_workloads.py $w 1 2>&1 | tail -1 >> $O/${T}_phases_$w.jsonl; done; done
