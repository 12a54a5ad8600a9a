#!/bin/bash
# round-2 trip H (8 GPUs): the driver's scaling sweep N = 1, 2, 4, 8 (C2, weak scaling) + C4 / C5 at N = 8
mkdir -p gpurun_out
run() { n=$1; shift; name=$1; shift
  if [ "$n" = "1" ]; then timeout 300 python bench.py --gpus 1 --steps 20 --warmup 5 --no-profile --no-cpu-baseline "$@" > gpurun_out/r2h_$name.json 2> gpurun_out/r2h_$name.err
  else timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 295$n$n bench.py --gpus $n --steps 20 --warmup 5 --no-profile "$@" > gpurun_out/r2h_$name.json 2> gpurun_out/r2h_$name.err; fi; }
run 1 scale1
run 2 scale2
run 4 scale4
run 8 scale8
run 8 scale8_after --allreduce after
run 8 scale8_eager --no-graph
run 8 scale8_c4 --config c4
run 8 scale8_c5 --config c5
for f in gpurun_out/r2h_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("step_mode","")[:60])
except Exception as e: print("ERR", e)
PY
done
tail -3 gpurun_out/r2h_scale8.err
