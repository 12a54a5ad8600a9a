#!/bin/bash
# round-2 trip M (gpurun --gpus 2): overlap actually enabled in the captured step (reducer state reset before capture)
mkdir -p gpurun_out
T0=$(date +%s)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 100 $TR --master-port 29563 tools/dp_equivalence.py --peer --graph > gpurun_out/r2m_equiv_peer_graph.log 2>&1; echo "rc=$?" >> gpurun_out/r2m_equiv_peer_graph.log
grep -E "^\{|rc=" gpurun_out/r2m_equiv_peer_graph.log | cut -c1-400
run() { name=$1; shift
  timeout 120 $TR --master-port 29565 bench.py --gpus 2 --steps 30 --warmup 5 --no-profile "$@" > gpurun_out/r2m_bench_$name.json 2> gpurun_out/r2m_bench_$name.err; echo "rc=$?" >> gpurun_out/r2m_bench_$name.err; }
run peer --allreduce peer
run peer_ov6 --allreduce peer --overlap-chunks 6
run peer_ov2 --allreduce peer --overlap-chunks 2
run peer_sm --allreduce peer --peer-ctas 0 --peer-tail-ctas -1
run peer_v1 --allreduce peer --peer-ctas 48 --peer-tail-ctas -1
for f in gpurun_out/r2m_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["gpu_launches"], d.get("gradient_exchange",{}).get("note"), d.get("invalid"))
except Exception as e: print("ERR", e)
PY
done
tail -3 gpurun_out/r2m_bench_peer.err | cut -c1-300
echo "total elapsed=$(( $(date +%s) - T0 ))s"
