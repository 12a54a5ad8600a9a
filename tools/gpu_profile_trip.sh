#!/bin/bash
# ncu evidence for profiles/: (1) launch list of one steady-state step, (2) DRAM traffic / tensor
# activity per launch (cheap metric set), (3) --set full on a few GEMM launches.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
CMD="python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-profile"
timeout -k 5 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum --clock-control none -s 1700 -c 700 --csv --log-file gpurun_out/step_metrics.csv $CMD > gpurun_out/ncu_step.log 2>&1
tail -1 gpurun_out/ncu_step.log | cut -c1-200; wc -l gpurun_out/step_metrics.csv
timeout -k 5 300 ncu --set full --clock-control none --import-source on -k regex:gemm -s 470 -c 6 -o gpurun_out/gemm_full -f $CMD > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log | cut -c1-200; ls -la gpurun_out/*.ncu-rep
