#!/bin/bash
# Round-1 validation trip: new-path tests first, then the whole gpu suite, then bench A/B and the
# ncu per-launch metrics of one steady-state step.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
timeout -k 10 420 python -m pytest tests/test_heads_gpu.py tests/test_optim_gpu.py tests/test_gemm_gpu.py \
    tests/test_model_gpu.py tests/test_rowops_gpu.py tests/test_attn_gpu.py -m gpu -q --timeout 120 \
    -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "exit=$? elapsed=$(( $(date +%s) - T0 ))s" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout -k 10 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit=$?"; cut -c1-400 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
UB200_GROUP_BN=128 timeout -k 10 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-profile > gpurun_out/bench_group128.json 2> gpurun_out/bench_group128.err
echo "bench group128 exit=$?"; cut -c1-250 gpurun_out/bench_group128.json
for v in 2 3; do
  UB200_LN_BWD_CTAS_PER_SM=$v timeout -k 10 150 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/bench_lnbwd$v.json 2> gpurun_out/bench_lnbwd$v.err
  echo "bench lnbwd$v exit=$?"; cut -c1-250 gpurun_out/bench_lnbwd$v.json
done
timeout -k 10 120 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit=$?"; tail -2 gpurun_out/smoke.log
if [ "${1:-}" = "ncu" ]; then
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile"
  timeout -k 5 300 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none -s 1500 -c 800 --csv --log-file gpurun_out/step_metrics.csv $CMD > gpurun_out/ncu_step.log 2>&1
  tail -1 gpurun_out/ncu_step.log | cut -c1-200; wc -l gpurun_out/step_metrics.csv
fi
echo "total elapsed=$(( $(date +%s) - T0 ))s"
