O=gpurun_out; T=${1:-r3c}; shift
cp surtr_b200/libsurtr_b200.so /tmp/lib_base.so
for v in base "$@"; do
  if [ $v = base ]; then cp /tmp/lib_base.so surtr_b200/libsurtr_b200.so; else cp surtr_b200/variants/lib_$v.so surtr_b200/libsurtr_b200.so; fi
  for rep in 1 2; do echo -n "$v " ; python tools/gpu_cold_latency.py 2>&1 | tail -1; done
done | tee $O/${T}_cold.txt
cp /tmp/lib_base.so surtr_b200/libsurtr_b200.so
