#!/bin/bash
# N-GPU variants of the gradient all-reduce overlap (run under: gpurun --gpus N -- bash tools/gpu_scale_trip.sh N)
N=${1:-2}
mkdir -p gpurun_out
port=29500
run() {
  tag=$1; shift
  port=$((port + 1))
  timeout -k 5 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port $port bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-profile "$@" \
    > gpurun_out/scale_${N}_${tag}.json 2> gpurun_out/scale_${N}_${tag}.err
  python - <<PY
import json
try:
    j = json.load(open("gpurun_out/scale_${N}_${tag}.json"))
    print("${tag}", j["value"], j["ms_per_step"], j["e2e"]["value"])
except Exception as e:
    print("${tag} FAILED", e)
PY
}
run ov4_r0 --overlap-chunks 4 --sm-reserve 0
run ov1 --overlap-chunks 1 --sm-reserve 0
run ov4_r16 --overlap-chunks 4 --sm-reserve 16
run ov4_r8 --overlap-chunks 4 --sm-reserve 8
run ov6_r16 --overlap-chunks 6 --sm-reserve 16
