#!/bin/bash
mkdir -p gpurun_out
python scripts/nvdec_probe.py > gpurun_out/r02_nvdec_probe.txt 2>&1
( time python -m pytest tests -m gpu -q ) > gpurun_out/r02_pytest_gpu.log 2>&1
tail -5 gpurun_out/r02_pytest_gpu.log
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none \
  -k regex:'temporal|act4|act_kernel|dst_|hough|ppht|noise|threshold|classic|preproc|stack|compact|replay|window|expand' -c 400 --csv \
  --log-file gpurun_out/r02_launches.csv \
  python bench.py --steps 2 --warmup 1 --no-e2e --no-next-rows --no-configs --no-extra --no-parity --no-cpu-baseline > gpurun_out/r02_launches_bench.log 2>&1
python scripts/t3_tune.py 3840 2160 30 512 > gpurun_out/r02_t3_variants_n30.txt 2>&1
python scripts/t3_tune.py 3840 2160 60 512 > gpurun_out/r02_t3_variants_n60.txt 2>&1
tail -3 gpurun_out/r02_t3_variants_n30.txt
cat gpurun_out/r02_nvdec_probe.txt
