O=gpurun_out; T=${1:-r2s}
python -m pytest tests -m gpu -x -q > $O/${T}_pytest.log 2>&1; tail -1 $O/${T}_pytest.log
python tests/measure/gpu_dofracture_trace.py > $O/${T}_dofracture.out 2> $O/${T}_dofracture_trace.txt
for w in config4 config3 config2 mesh; do EVENTS=256 python tools/gpu_profile_workloads.py $w 1 2>&1 | tail -1 >> $O/${T}_phases_$w.jsonl; done
