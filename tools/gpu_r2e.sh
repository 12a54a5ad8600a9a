#!/bin/bash
# round-2 trip E: A/B of the single-block attention kernels, split LN backward, fused LN; C4/C5 lines
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attn_gpu.py tests/test_rowops_gpu.py tests/test_gemm_gpu.py -x -q -m gpu 2>&1 | tail -8 > gpurun_out/r2e_kernel_tests.log
for cfg in "0 0" "1 0" "0 1" "1 1"; do
  set -- $cfg
  UB200_FUSE_LN=$1 UB200_LN_BWD_SPLIT=$2 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-profile > gpurun_out/r2e_bench_fuse$1_split$2.json 2> gpurun_out/r2e_bench_fuse$1_split$2.err
done
UB200_FUSE_LN=1 UB200_LN_BWD_SPLIT=1 timeout 400 python bench.py --config c4 --steps 24 --warmup 4 --no-profile > gpurun_out/r2e_bench_c4_fuse1_split1.json 2> gpurun_out/r2e_bench_c4.err
UB200_FUSE_LN=0 UB200_LN_BWD_SPLIT=1 timeout 400 python bench.py --config c4 --steps 24 --warmup 4 --no-profile > gpurun_out/r2e_bench_c4_fuse0_split1.json 2>> gpurun_out/r2e_bench_c4.err
UB200_FUSE_LN=0 UB200_LN_BWD_SPLIT=0 timeout 400 python bench.py --config c4 --steps 24 --warmup 4 --no-profile > gpurun_out/r2e_bench_c4_fuse0_split0.json 2>> gpurun_out/r2e_bench_c4.err
timeout 300 python bench.py --config c5 --steps 16 --warmup 3 > gpurun_out/r2e_bench_c5.json 2> gpurun_out/r2e_bench_c5.err
UB200_FUSE_LN=1 UB200_LN_BWD_SPLIT=1 timeout 600 python -m pytest tests/test_model_gpu.py tests/test_c2_parity_gpu.py tests/test_graphed_gpu.py -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2e_model_tests_fused.log
tail -5 gpurun_out/r2e_kernel_tests.log; tail -5 gpurun_out/r2e_model_tests_fused.log
for f in gpurun_out/r2e_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"])
except Exception as e: print("ERR", e)
PY
done
tail -3 gpurun_out/r2e_bench_fuse1_split1.err gpurun_out/r2e_bench_c4.err
