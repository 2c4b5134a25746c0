#!/bin/bash
# first-contact GPU check: build is in-tree already; run the gpu tests verbosely
set -x
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
mkdir -p gpurun_out
python -m pytest tests -m gpu -q -s 2>&1 | tail -60 > gpurun_out/pytest_gpu.log
tail -60 gpurun_out/pytest_gpu.log
