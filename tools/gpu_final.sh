#!/bin/bash
# Round-end evidence trip: gpu suite, default bench (both arms), ncu launch list of the bench
# command, ncu --set full of the top kernels (CSV pages exported on the box; the .ncu-rep stays there).
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
timeout -k 10 420 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "exit=$? elapsed=$(( $(date +%s) - T0 ))s" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|passed|failed|exit=" gpurun_out/pytest_gpu.log | tail -8
timeout -k 10 300 python bench.py > gpurun_out/bench_final.json 2> gpurun_out/bench_final.err
echo "bench exit=$? t=$(( $(date +%s) - T0 ))s"; cut -c1-260 gpurun_out/bench_final.json; tail -2 gpurun_out/bench_final.err
timeout -k 10 200 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_reference.json 2> gpurun_out/bench_reference.err
echo "reference arm exit=$? t=$(( $(date +%s) - T0 ))s"; cut -c1-300 gpurun_out/bench_reference.json
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile"
timeout -k 5 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none -s 1400 -c 800 --csv --log-file gpurun_out/step_metrics.csv $CMD > gpurun_out/ncu_step.log 2>&1
wc -l gpurun_out/step_metrics.csv; echo "launch list t=$(( $(date +%s) - T0 ))s"
timeout -k 5 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:gemm_group_kernel|gemm_kernel<.int.192, .bool.0, .bool.0, .bool.1, .int.9>|gemm2sm_kernel<.int.256, .bool.0, .bool.0, .bool.1, .int.1>|attn_bwd_kernel|ln_bwd_kernel<.bool.1, .int.3>' \
    -s 200 -c 8 -o /tmp/top_full -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-160
ncu -i /tmp/top_full.ncu-rep --page raw --csv > gpurun_out/top_full_raw.csv 2>/dev/null
ncu -i /tmp/top_full.ncu-rep --page details --csv > gpurun_out/top_full_details.csv 2>/dev/null
ls -la gpurun_out/top_full_*.csv /tmp/top_full.ncu-rep
echo "total elapsed=$(( $(date +%s) - T0 ))s"
