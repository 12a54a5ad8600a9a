#!/bin/bash
# round-2 trip F (gpurun --gpus 2): NCCL equivalence (eager + in-graph all-reduce), 2-GPU bench variants
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_dist_nccl2_gpu.py tests/test_dist_gpu.py -q -m gpu 2>&1 | tail -15 > gpurun_out/r2f_nccl_tests.log
run() { # name, extra args
  name=$1; shift
  timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 30 --warmup 5 --no-profile "$@" > gpurun_out/r2f_bench2_$name.json 2> gpurun_out/r2f_bench2_$name.err
}
run ingraph --allreduce in-graph
run after --allreduce after
run eager --no-graph
run ingraph_ov1 --allreduce in-graph --overlap-chunks 1
timeout 300 python bench.py --gpus 1 --steps 30 --warmup 5 --no-profile --no-cpu-baseline > gpurun_out/r2f_bench1.json 2> gpurun_out/r2f_bench1.err
tail -8 gpurun_out/r2f_nccl_tests.log
for f in gpurun_out/r2f_bench*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["step_mode"][:70])
except Exception as e: print("ERR", e)
PY
done
tail -4 gpurun_out/r2f_bench2_ingraph.err
