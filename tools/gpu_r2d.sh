#!/bin/bash
# round-2 trip D: fused residual+LayerNorm GEMM epilogue (unit test, A/B bench), C5, sanitizer, full GPU suite
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gemm_gpu.py -k fused -x -q -m gpu 2>&1 | tail -12 > gpurun_out/r2d_fused_ln_test.log
if grep -q "passed" gpurun_out/r2d_fused_ln_test.log && ! grep -q "failed" gpurun_out/r2d_fused_ln_test.log; then
  UB200_FUSE_LN=1 timeout 400 python -m pytest tests/test_model_gpu.py tests/test_c2_parity_gpu.py -x -q -m gpu 2>&1 | tail -6 > gpurun_out/r2d_fused_ln_model.log
  UB200_FUSE_LN=1 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2d_bench_fuse1.json 2> gpurun_out/r2d_bench_fuse1.err
fi
UB200_FUSE_LN=0 timeout 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline --no-profile > gpurun_out/r2d_bench_fuse0.json 2> gpurun_out/r2d_bench_fuse0.err
timeout 300 python bench.py --config c5 --steps 16 --warmup 3 > gpurun_out/r2d_bench_c5.json 2> gpurun_out/r2d_bench_c5.err
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 python tools/sanitize_kernels.py > gpurun_out/r2d_memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2d_memcheck.log
timeout 600 compute-sanitizer --tool racecheck --error-exitcode 7 python tools/sanitize_kernels.py > gpurun_out/r2d_racecheck.log 2>&1; echo "racecheck rc=$?" >> gpurun_out/r2d_racecheck.log
python -m pytest tests -q -m gpu 2>&1 | tail -15 > gpurun_out/r2d_tests_all.log
for f in r2d_fused_ln_test.log r2d_fused_ln_model.log r2d_tests_all.log; do echo "== $f"; tail -8 gpurun_out/$f; done
for f in r2d_bench_fuse1 r2d_bench_fuse0 r2d_bench_c5; do echo "== $f"; cut -c1-900 gpurun_out/$f.json; tail -3 gpurun_out/$f.err; done
tail -6 gpurun_out/r2d_memcheck.log; tail -6 gpurun_out/r2d_racecheck.log
