#!/bin/bash
# round-2 trip C: new tests, C2 + C4 bench (graph mode), ncu launch list of the graph-mode step
mkdir -p gpurun_out
python -m pytest tests/test_graphed_gpu.py tests/test_reference_heads_gpu.py tests/test_optim_gpu.py tests/test_heads_gpu.py -x -q -m gpu 2>&1 | tail -25 > gpurun_out/r2c_tests.log
timeout 600 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2c_bench_c2.json 2> gpurun_out/r2c_bench_c2.err
timeout 900 python bench.py --config c4 --steps 24 --warmup 4 > gpurun_out/r2c_bench_c4.json 2> gpurun_out/r2c_bench_c4.err
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 3000 -c 700 --csv --log-file gpurun_out/r2c_launches.csv python bench.py --steps 3 --warmup 2 --no-cpu-baseline --no-profile --no-graph > gpurun_out/r2c_ncu.log 2>&1
tail -12 gpurun_out/r2c_tests.log; cat gpurun_out/r2c_bench_c2.json | cut -c1-1500; tail -3 gpurun_out/r2c_bench_c2.err; cat gpurun_out/r2c_bench_c4.json | cut -c1-2500; tail -5 gpurun_out/r2c_bench_c4.err; tail -3 gpurun_out/r2c_ncu.log; wc -l gpurun_out/r2c_launches.csv
