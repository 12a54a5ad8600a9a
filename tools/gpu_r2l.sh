#!/bin/bash
# round-2 trip L (gpurun --gpus N, default 2): peer exchange, copy-engine form (memcpy nodes + flag kernels + local
# reduction), equal length profile on every rank.
N=${1:-2}
mkdir -p gpurun_out
T0=$(date +%s)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1"
timeout 120 $TR --master-port 29561 tools/peer_check.py > gpurun_out/r2l_n${N}_peer_check.json 2> gpurun_out/r2l_n${N}_peer_check.err; echo "rc=$?" >> gpurun_out/r2l_n${N}_peer_check.err
echo "peer_check t=$(( $(date +%s) - T0 ))s"; cut -c1-1100 gpurun_out/r2l_n${N}_peer_check.json; tail -3 gpurun_out/r2l_n${N}_peer_check.err | cut -c1-300
timeout 100 $TR --master-port 29563 tools/dp_equivalence.py --peer --graph > gpurun_out/r2l_n${N}_equiv_peer_graph.log 2>&1; echo "rc=$?" >> gpurun_out/r2l_n${N}_equiv_peer_graph.log
grep -E "^\{|rc=" gpurun_out/r2l_n${N}_equiv_peer_graph.log | cut -c1-400
echo "equiv t=$(( $(date +%s) - T0 ))s"
run() { name=$1; shift
  timeout 120 $TR --master-port 29565 bench.py --gpus $N --steps 30 --warmup 5 --no-profile "$@" > gpurun_out/r2l_n${N}_bench_$name.json 2> gpurun_out/r2l_n${N}_bench_$name.err; echo "rc=$?" >> gpurun_out/r2l_n${N}_bench_$name.err; }
run peer --allreduce peer
if [ "$N" = "2" ]; then
run peer_ov1 --allreduce peer --overlap-chunks 1
run peer_ov6 --allreduce peer --overlap-chunks 6
run peer_sm --allreduce peer --peer-ctas 0 --peer-tail-ctas 0
fi
run after --allreduce after
for f in gpurun_out/r2l_n${N}_bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["gpu_launches"], d["step_mode"][:60], d.get("gradient_exchange",{}).get("note"), d.get("invalid"))
except Exception as e: print("ERR", e)
PY
done
tail -3 gpurun_out/r2l_n${N}_bench_peer.err | cut -c1-300
echo "total elapsed=$(( $(date +%s) - T0 ))s"
