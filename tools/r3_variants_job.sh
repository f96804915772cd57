# A/B of prebuilt library variants (surtr_b200/variants/lib_<name>.so, built with -D switches): phase times per variant
O=gpurun_out; T=${1:-r3v}; shift
cp surtr_b200/libsurtr_b200.so /tmp/lib_base.so
for v in base "$@"; do
  if [ $v = base ]; then cp /tmp/lib_base.so surtr_b200/libsurtr_b200.so; else cp surtr_b200/variants/lib_$v.so surtr_b200/libsurtr_b200.so; fi
  for rep in 1 2; do for w in config4 config3 config2; do
    EVENTS=256 python tools/gpu_profile_workloads.py $w 1 2>&1 | tail -1 >> $O/${T}_${v}_${w}.jsonl
  done; done
done
cp /tmp/lib_base.so surtr_b200/libsurtr_b200.so
python - <<P
import json,glob
for f in sorted(glob.glob('gpurun_out/${T}_*.jsonl')):
    for l in open(f):
        try: d=json.loads(l)
        except Exception: print(f, l[:200]); continue
        k=d['kernel_ms_unprofiled']
        print(f.split('${T}_')[1], d['fragments'], 'k3', round(k['k3_clip_small'],4), 'k3large', round(k['k3_clip_large'],4), 'k4', round(k['k4_gather'],4), 'sum', round(sum(k.values()),4))
P
