#!/bin/bash
# Short trip: whole gpu suite + one bench + smoke.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
timeout -k 10 420 python -m pytest tests -m gpu -q --timeout 120 -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "exit=$? elapsed=$(( $(date +%s) - T0 ))s" >> gpurun_out/pytest_gpu.log
grep -E "^FAILED|passed|failed|exit=" gpurun_out/pytest_gpu.log | tail -15
timeout -k 10 150 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err
echo "bench default exit=$?"; cut -c1-300 gpurun_out/bench_default.json; tail -3 gpurun_out/bench_default.err
for extra in "$@"; do
  env $extra timeout -k 10 150 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-profile > "gpurun_out/bench_${extra//[^A-Za-z0-9_=]/_}.json" 2>/dev/null
  echo "bench $extra exit=$?"; cut -c1-200 "gpurun_out/bench_${extra//[^A-Za-z0-9_=]/_}.json"
done
echo "total elapsed=$(( $(date +%s) - T0 ))s"
if [ -n "${NCU_FULL:-}" ]; then
  CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile"
  timeout -k 5 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
      -k "regex:$NCU_FULL" -s ${NCU_SKIP:-40} -c ${NCU_COUNT:-4} -o gpurun_out/gemm_epi_full -f $CMD > gpurun_out/ncu_full.log 2>&1
  tail -2 gpurun_out/ncu_full.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep
  echo "total elapsed=$(( $(date +%s) - T0 ))s"
fi
