#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/r02_topo_n8.txt 2>&1
lscpu | grep -i "numa\|socket\|model name\|^CPU(s)" > gpurun_out/r02_lscpu_n8.txt 2>&1
BENCH_TRACE=1 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29533 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/r02_bench_n8.json 2> gpurun_out/r02_bench_n8.err
python - <<PY
import json
d=json.loads(open("gpurun_out/r02_bench_n8.json").read().strip().splitlines()[-1])
print("value", d["value"], "ms/step", d["ms_per_step"], "parity", d["sharded_parity"], "e2e", d["e2e"])
PY
grep "step times\|phases" gpurun_out/r02_bench_n8.err | head -20
