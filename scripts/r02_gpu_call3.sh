#!/bin/bash
mkdir -p gpurun_out
( time python -m pytest tests -m gpu -q -x -k "per_frame or temporal3 or classic or error_behaviour or mixed or sliding_window or longer_than or config1 or batched_api" ) > gpurun_out/r02_pytest_gpu_b.log 2>&1
tail -5 gpurun_out/r02_pytest_gpu_b.log
python scripts/per_frame_rate.py 30 > gpurun_out/r02_per_frame_rate.txt 2>&1
cat gpurun_out/r02_per_frame_rate.txt
python scripts/chain_run.py 3840 2160 60 512 1 1 5 > gpurun_out/r02_chain_cfg4_maskw.txt 2>&1
tail -2 gpurun_out/r02_chain_cfg4_maskw.txt
python scripts/chain_run.py 3840 2160 30 512 1 1 5 | tail -1
