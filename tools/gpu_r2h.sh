#!/bin/bash
# round-2 trip H (gpurun --gpus 8): N = 8 weak scaling of C2 in the all-reduce modes (+ C4), tight timeouts
mkdir -p gpurun_out
run() { name=$1; shift
  timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29588 bench.py --gpus 8 --steps 20 --warmup 5 --no-profile "$@" > gpurun_out/r2h_$name.json 2> gpurun_out/r2h_$name.err; echo "rc=$?" >> gpurun_out/r2h_$name.err; }
run scale8_after --allreduce after
run scale8_split --allreduce split
run scale8_c4 --config c4 --allreduce split
for f in gpurun_out/r2h_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(d["n_gpus"], d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d.get("step_mode","")[:80])
except Exception as e: print("ERR", e)
PY
done
tail -3 gpurun_out/r2h_scale8_split.err
