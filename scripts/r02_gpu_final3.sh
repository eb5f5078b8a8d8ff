#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -3 gpurun_out/r02_smoke.log
( time python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err ) 2>&1 | tail -3
