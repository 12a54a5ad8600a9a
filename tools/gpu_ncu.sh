#!/bin/bash
# ncu evidence trip: (A) --set full + source for the epilogue-bound GEMMs, (B) section summary of
# every library kernel of one steady-state step.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile"
timeout -k 5 240 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:gemm_kernel<.int.192, .bool.0, .bool.[01], .bool.1, .int.(9|144)>' -s 48 -c 4 \
    -o gpurun_out/gemm_epi_full -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -2 gpurun_out/ncu_full.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep
echo "A elapsed=$(( $(date +%s) - T0 ))s"
timeout -k 5 400 ncu --section SpeedOfLight --section MemoryWorkloadAnalysis --section LaunchStats \
    --section Occupancy --section WarpStateStats --clock-control none -k 'regex:^ub::|ub::' -s 1130 -c 230 \
    -o gpurun_out/step_sections -f $CMD > gpurun_out/ncu_step_sections.log 2>&1
tail -2 gpurun_out/ncu_step_sections.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep
echo "total elapsed=$(( $(date +%s) - T0 ))s"
