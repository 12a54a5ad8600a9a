#!/bin/bash
# GPU bring-up trip: run the gpu-marked tests (verbose, keep going) and save the log.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
timeout ${1:-900} python -m pytest tests -m gpu -q -x --timeout 300 -p no:cacheprovider ${2:-} > gpurun_out/pytest_gpu.log 2>&1
echo "exit=$?" >> gpurun_out/pytest_gpu.log
tail -40 gpurun_out/pytest_gpu.log
