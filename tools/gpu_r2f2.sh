#!/bin/bash
# round-2 trip F2 (gpurun --gpus 2): split-graph overlap mode, TIGHT timeouts (a hang costs 2x GPU minutes)
mkdir -p gpurun_out
timeout 150 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29541 tools/dp_equivalence.py --graph > gpurun_out/r2f2_equiv_graph.log 2>&1; echo "rc=$?" >> gpurun_out/r2f2_equiv_graph.log
run() { name=$1; shift
  timeout 170 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29543 bench.py --gpus 2 --steps 30 --warmup 5 --no-profile "$@" > gpurun_out/r2f2_bench2_$name.json 2> gpurun_out/r2f2_bench2_$name.err; echo "rc=$?" >> gpurun_out/r2f2_bench2_$name.err; }
run split --allreduce split
run split_ov2 --allreduce split --overlap-chunks 2
run split_ov6 --allreduce split --overlap-chunks 6
run after --allreduce after
grep -v "^$" gpurun_out/r2f2_equiv_graph.log | tail -4
for f in gpurun_out/r2f2_bench2_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["step_mode"][:90])
except Exception as e: print("ERR", e)
PY
done
tail -5 gpurun_out/r2f2_bench2_split.err
