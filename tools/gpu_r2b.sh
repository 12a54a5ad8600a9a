#!/bin/bash
# round-2 trip B: full GPU suite after the arena / graph refactor + bench (graph and eager)
mkdir -p gpurun_out
python -m pytest tests -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2b_tests.log
python -m pytest tests/test_c2_parity_gpu.py tests/test_reference_heads_gpu.py tests/test_graphed_gpu.py -q -m gpu -s 2>&1 | grep -v "^$" | tail -40 > gpurun_out/r2b_tests2.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2b_bench_graph.json 2> gpurun_out/r2b_bench_graph.err
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-graph --no-profile > gpurun_out/r2b_bench_eager.json 2> gpurun_out/r2b_bench_eager.err
tail -12 gpurun_out/r2b_tests.log; tail -30 gpurun_out/r2b_tests2.log; cat gpurun_out/r2b_bench_graph.json; tail -5 gpurun_out/r2b_bench_graph.err; cat gpurun_out/r2b_bench_eager.json; tail -5 gpurun_out/r2b_bench_eager.err
