#!/bin/bash
# N-GPU sanity of the driver's launch line (gpurun --gpus N -- bash tools/gpu_scale8.sh N)
N=${1:-8}
mkdir -p gpurun_out
for v in "ov4 --overlap-chunks 4" "ov1 --overlap-chunks 1"; do
  set -- $v; tag=$1; shift
  timeout -k 5 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 \
    --master-port 29511 bench.py --gpus $N --steps 20 --warmup 5 --no-cpu-baseline --no-profile "$@" \
    > gpurun_out/scale_${N}_${tag}.json 2> gpurun_out/scale_${N}_${tag}.err
  echo "rc=$? lines=$(wc -l < gpurun_out/scale_${N}_${tag}.json)"
  python -c "
import json; j=json.load(open('gpurun_out/scale_${N}_${tag}.json')); print('${tag}', j['value'], j['ms_per_step'], j['e2e']['value'], j['clocks'])" || tail -5 gpurun_out/scale_${N}_${tag}.err
done
