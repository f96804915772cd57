O=gpurun_out; T=${1:-r2san}
SEL="golden_fixture or edge_cases or degenerate or one_copy or packed_wire or large_tier or global_tier_mesh or high_valence or failed_pairs or resident_pattern or config4_batched"
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "$SEL" > $O/${T}_memcheck.log 2>&1; echo memcheck rc=$? >> $O/${T}_memcheck.log
timeout 400 compute-sanitizer --tool synccheck --error-exitcode 1 python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_fixture or degenerate or one_copy or large_tier or global_tier_mesh" > $O/${T}_synccheck.log 2>&1; echo synccheck rc=$? >> $O/${T}_synccheck.log
timeout 400 compute-sanitizer --tool racecheck --racecheck-report analysis python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden_fixture or degenerate_cuts or one_copy" > $O/${T}_racecheck.log 2>&1; echo racecheck rc=$? >> $O/${T}_racecheck.log
# from-scratch build on the box + smoke
python -c "
import time, __graft_entry__ as g
t=time.time(); g.build_cuda(force=True); g.build_host(force=True); print('built from scratch on the GPU box in %.1f s' % (time.time()-t)); g.smoke(); print('smoke ok')
" > $O/${T}_box_build.txt 2>&1
