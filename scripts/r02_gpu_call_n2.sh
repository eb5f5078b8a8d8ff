#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n2.txt 2>&1
( time python -m pytest tests/test_gpu_multi.py -m gpu -q ) > gpurun_out/r02_pytest_gpu_multi.log 2>&1
tail -4 gpurun_out/r02_pytest_gpu_multi.log
BENCH_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r02_bench_n2.json 2> gpurun_out/r02_bench_n2.err
tail -c 1500 gpurun_out/r02_bench_n2.json; tail -5 gpurun_out/r02_bench_n2.err
