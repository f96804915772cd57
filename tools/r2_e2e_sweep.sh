O=gpurun_out; T=r2x
python bench.py --no-cpu-baseline --no-secondary > $O/${T}_bench_b64.json 2> $O/${T}_bench_b64.err
python bench.py --no-cpu-baseline --no-secondary --e2e-batch 128 > $O/${T}_bench_b128.json 2> $O/${T}_bench_b128.err
python bench.py --no-cpu-baseline --no-secondary --e2e-batch 128 --e2e-contexts 4 > $O/${T}_bench_b128c4.json 2> $O/${T}_bench_b128c4.err
python bench.py --no-cpu-baseline --no-secondary --e2e-batch 32 --e2e-contexts 4 > $O/${T}_bench_b32c4.json 2> $O/${T}_bench_b32c4.err
