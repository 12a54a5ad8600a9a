#!/bin/bash
# round-2 trip G (1 GPU): what the driver runs at round end + fp16 line + ncu launch list / top-kernel capture
mkdir -p gpurun_out
python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/r2g_tests_all.log
python __graft_entry__.py smoke > gpurun_out/r2g_smoke.log 2>&1
timeout 400 python bench.py --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_bench_default.json 2> gpurun_out/r2g_bench_default.err
timeout 300 python bench.py --impl reference --gpus 1 --steps 20 --warmup 5 > gpurun_out/r2g_bench_reference.json 2> gpurun_out/r2g_bench_reference.err
timeout 300 python bench.py --dtype fp16 --steps 30 --warmup 5 --no-cpu-baseline --no-profile > gpurun_out/r2g_bench_fp16.json 2> gpurun_out/r2g_bench_fp16.err
timeout 900 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none -s 700 -c 560 --csv --log-file gpurun_out/r2g_step_metrics.csv python bench.py --steps 2 --warmup 1 --no-graph --no-profile --no-cpu-baseline > gpurun_out/r2g_ncu.log 2>&1
tail -6 gpurun_out/r2g_tests_all.log; tail -2 gpurun_out/r2g_smoke.log
for f in default reference fp16; do echo "== $f"; cut -c1-1200 gpurun_out/r2g_bench_$f.json; tail -2 gpurun_out/r2g_bench_$f.err; done
wc -l gpurun_out/r2g_step_metrics.csv
