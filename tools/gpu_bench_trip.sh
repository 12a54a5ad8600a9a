#!/bin/bash
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
echo "=== smoke"; timeout 300 python __graft_entry__.py smoke 2>&1 | tail -5
echo "=== model tests"; timeout 300 python -m pytest tests/test_model_gpu.py -m gpu -q -p no:cacheprovider 2>&1 | tail -5
echo "=== bench"; timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/bench.json 2> gpurun_out/bench.err; tail -3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ "$1" == "ncu" ]; then
echo "=== ncu launch list"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 2000 -c 700 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-profile > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log; wc -l gpurun_out/launches.csv
fi
