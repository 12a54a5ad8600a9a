#!/bin/bash
# round-2 trip N (1 GPU): what the driver runs at round end — gpu suite, smoke, default bench (both arms) + fp16 line
mkdir -p gpurun_out
T0=$(date +%s)
timeout 420 python -m pytest tests -q -m gpu -x 2>&1 | tail -8 > gpurun_out/r2n_tests.log; echo "tests t=$(( $(date +%s) - T0 ))s"; tail -4 gpurun_out/r2n_tests.log
timeout 120 python __graft_entry__.py smoke > gpurun_out/r2n_smoke.log 2>&1; tail -1 gpurun_out/r2n_smoke.log
timeout 300 python bench.py > gpurun_out/r2n_bench_default.json 2> gpurun_out/r2n_bench_default.err; echo "bench rc=$? t=$(( $(date +%s) - T0 ))s"
timeout 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2n_bench_reference.json 2> gpurun_out/r2n_bench_reference.err; echo "ref rc=$? t=$(( $(date +%s) - T0 ))s"
timeout 200 python bench.py --dtype fp16 --steps 30 --warmup 5 --no-cpu-baseline --no-profile > gpurun_out/r2n_bench_fp16.json 2> gpurun_out/r2n_bench_fp16.err; echo "fp16 rc=$?"
for f in default reference fp16; do echo "== $f"; cut -c1-700 gpurun_out/r2n_bench_$f.json; tail -2 gpurun_out/r2n_bench_$f.err | cut -c1-300; done
echo "total elapsed=$(( $(date +%s) - T0 ))s"
