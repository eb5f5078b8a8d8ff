#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -3 gpurun_out/r02_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_smoke.log 2>&1; tail -3 gpurun_out/r02_smoke.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k regex:'temporal|act4|act_kernel|dst_|hough|ppht|noise|threshold|classic|preproc|stack|compact|replay|window|expand|pf_|mfnr' -c 400 --csv \
  --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-next-rows --no-configs --no-extra --no-parity --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
( time python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err ) 2>&1 | tail -3
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_ref.json 2> gpurun_out/r02_bench_ref.err
( time timeout 600 compute-sanitizer --tool memcheck python scripts/sanitize_case.py ) > gpurun_out/r02_sanitizer_memcheck.log 2>&1
tail -2 gpurun_out/r02_sanitizer_memcheck.log
( time timeout 900 compute-sanitizer --tool racecheck python scripts/sanitize_case.py ) > gpurun_out/r02_sanitizer_racecheck.log 2>&1
tail -2 gpurun_out/r02_sanitizer_racecheck.log
