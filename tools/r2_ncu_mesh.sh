O=gpurun_out; T=r2v
ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:clip_global_kernel -c 1 -f -o $O/${T}_mesh_k3 python tools/gpu_profile_workloads.py mesh 1 > $O/${T}_mesh_k3.log 2>&1
python tools/gpu_profile_workloads.py mesh 1 2>&1 | tail -1 > $O/${T}_phases_mesh.json
