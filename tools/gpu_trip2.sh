#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== gemm check"; timeout 400 python tools/gpu_gemm_check.py > gpurun_out/gemm_check.log 2>&1; grep -E "FAIL|time|epilogue|ALL_OK|SOME|Error|error" gpurun_out/gemm_check.log | head -40
echo "=== tests"; timeout 600 python -m pytest tests -m gpu -q -p no:cacheprovider -x 2>&1 | tail -8
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
