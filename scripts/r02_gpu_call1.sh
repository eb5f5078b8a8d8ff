#!/bin/bash
# round-2 evidence call: GPU tests, chain captures (traffic), launch list, bench line, sanitizers
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/r02_gpu.txt
ls /usr/lib/x86_64-linux-gnu/ | grep -i "nvcuvid\|nvidia-encode\|libcuda" >> gpurun_out/r02_gpu.txt 2>&1
( time python -m pytest tests -m gpu -x -q ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu.log
SOURCE_PAGE=1 scripts/ncu_chain.sh cfg3 3840 2160 30 512 1 0 5
scripts/ncu_chain.sh cfg2 1920 1080 5 2048 0 0 5
scripts/ncu_chain.sh cfg4 3840 2160 60 512 1 1 5
scripts/ncu_chain.sh cfg5 7680 4320 30 128 1 0 5
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-next-rows --no-configs --no-extra --no-parity --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err
tail -c 600 gpurun_out/r02_bench_n1.json
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
( time timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_case.py ) > gpurun_out/r02_sanitizer_memcheck.log 2>&1
tail -3 gpurun_out/r02_sanitizer_memcheck.log
( time timeout 700 compute-sanitizer --tool racecheck python scripts/sanitize_case.py ) > gpurun_out/r02_sanitizer_racecheck.log 2>&1
tail -3 gpurun_out/r02_sanitizer_racecheck.log
