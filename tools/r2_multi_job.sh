O=gpurun_out; T=${1:-r2n8}
nvidia-smi topo -m > $O/${T}_topo.txt 2>&1
python -m pytest tests/test_gpu_multi.py -m gpu -x -q > $O/${T}_pytest_multi.log 2>&1; tail -1 $O/${T}_pytest_multi.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 10 --warmup 3 > $O/${T}_bench_n8.json 2> $O/${T}_bench_n8.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 4 --steps 10 --warmup 3 --no-secondary > $O/${T}_bench_n4.json 2> $O/${T}_bench_n4.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --no-secondary > $O/${T}_bench_n2.json 2> $O/${T}_bench_n2.err
python bench.py --no-secondary --no-cpu-baseline > $O/${T}_bench_n1.json 2> $O/${T}_bench_n1.err
