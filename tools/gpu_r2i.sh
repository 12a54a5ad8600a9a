#!/bin/bash
# round-2 trip I (1 GPU): state check after the container restore + evidence for the next kernel work:
#  (a) default bench line (graph mode, event pass), (b) ncu launch list of one eager C2 step with
#  DRAM / L2 / tensor-pipe metrics, (c) ncu --set full + source of the non-GEMM kernels of one step
#  (attention, LayerNorm, embedding / head row kernels), raw + source pages exported on the box.
mkdir -p gpurun_out
export PYTHONUNBUFFERED=1
T0=$(date +%s)
timeout -k 10 300 python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/r2i_bench.json 2> gpurun_out/r2i_bench.err
echo "bench rc=$? t=$(( $(date +%s) - T0 ))s"; cut -c1-400 gpurun_out/r2i_bench.json; tail -2 gpurun_out/r2i_bench.err
CMD="python bench.py --steps 2 --warmup 1 --no-graph --no-profile --no-cpu-baseline"
timeout -k 5 400 ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,lts__t_bytes.sum \
    --clock-control none -s 700 -c 560 --csv --log-file gpurun_out/r2i_step_metrics.csv $CMD > gpurun_out/r2i_ncu_list.log 2>&1
wc -l gpurun_out/r2i_step_metrics.csv; echo "launch list t=$(( $(date +%s) - T0 ))s"
# non-GEMM library kernels of ~one step (12 layers x (attn_fwd, 2 ln_fwd, 2 ln_bwd, attn_bwd) + embedding / head)
timeout -k 5 420 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:ub::(attn_|ln_|embed_|wcolsum|ce_|gather_rows|cvt_|colsum|dgelu|add16)' -s 130 -c 110 \
    -o /tmp/r2i_rows -f $CMD > gpurun_out/r2i_ncu_rows.log 2>&1
tail -1 gpurun_out/r2i_ncu_rows.log | cut -c1-160; echo "rows full t=$(( $(date +%s) - T0 ))s"
ncu -i /tmp/r2i_rows.ncu-rep --page raw --csv > gpurun_out/r2i_rows_raw.csv 2>/dev/null
for k in attn_fwd_kernel attn_bwd_kernel ln_bwd_kernel ln_fwd_kernel embed_rows_fwd_kernel embed_bwd_scatter_kernel wcolsum_kernel; do
  ncu -i /tmp/r2i_rows.ncu-rep --page source --csv --kernel-name-base demangled -k regex:$k -c 1 > gpurun_out/r2i_src_$k.csv 2>/dev/null
done
ls -la /tmp/r2i_rows.ncu-rep gpurun_out/ | head -30
# one layer's worth of GEMMs, forward + backward roles (raw page only)
timeout -k 5 300 ncu --set full --clock-control none --import-source on --kernel-name-base demangled \
    -k 'regex:ub::gemm' -s 170 -c 16 -o /tmp/r2i_gemm -f $CMD > gpurun_out/r2i_ncu_gemm.log 2>&1
ncu -i /tmp/r2i_gemm.ncu-rep --page raw --csv > gpurun_out/r2i_gemm_raw.csv 2>/dev/null
for k in 'gemm_kernel' 'gemm2sm_kernel'; do
  ncu -i /tmp/r2i_gemm.ncu-rep --page source --csv --kernel-name-base demangled -k regex:ub::$k -c 1 > gpurun_out/r2i_src_$k.csv 2>/dev/null
done
echo "total elapsed=$(( $(date +%s) - T0 ))s"; du -sh gpurun_out
