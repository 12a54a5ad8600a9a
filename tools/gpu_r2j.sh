#!/bin/bash
# round-2 trip J (gpurun --gpus 2): the NVLink peer-memory gradient exchange (csrc/peer.cu) — correctness and
# bandwidth vs NCCL (tools/peer_check.py), 2-rank == 1-rank-with-accumulation (eager + in-graph), then the
# C2 bench at N = 2 in the exchange modes.  TIGHT timeouts (a hang costs 2x GPU minutes).
mkdir -p gpurun_out
T0=$(date +%s)
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1"
timeout 150 $TR --master-port 29561 tools/peer_check.py > gpurun_out/r2j_peer_check.json 2> gpurun_out/r2j_peer_check.err; echo "rc=$?" >> gpurun_out/r2j_peer_check.err
echo "peer_check t=$(( $(date +%s) - T0 ))s"; cut -c1-900 gpurun_out/r2j_peer_check.json; tail -4 gpurun_out/r2j_peer_check.err
timeout 120 $TR --master-port 29562 tools/dp_equivalence.py --peer > gpurun_out/r2j_equiv_peer.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_equiv_peer.log
timeout 120 $TR --master-port 29563 tools/dp_equivalence.py --peer --graph > gpurun_out/r2j_equiv_peer_graph.log 2>&1; echo "rc=$?" >> gpurun_out/r2j_equiv_peer_graph.log
grep -E "^\{|rc=" gpurun_out/r2j_equiv_peer.log gpurun_out/r2j_equiv_peer_graph.log | cut -c1-400
echo "equiv t=$(( $(date +%s) - T0 ))s"
run() { name=$1; shift
  timeout 150 $TR --master-port 29565 bench.py --gpus 2 --steps 30 --warmup 5 --no-profile "$@" > gpurun_out/r2j_bench2_$name.json 2> gpurun_out/r2j_bench2_$name.err; echo "rc=$?" >> gpurun_out/r2j_bench2_$name.err; }
run peer --allreduce peer
run peer_c16 --allreduce peer --peer-ctas 16 --peer-tail-ctas 32
run peer_c64 --allreduce peer --peer-ctas 64 --peer-tail-ctas 148
run peer_ov2 --allreduce peer --overlap-chunks 2
run peer_ov6 --allreduce peer --overlap-chunks 6
run after --allreduce after
run split --allreduce split
for f in gpurun_out/r2j_bench2_*.json; do echo "== $f"; python - "$f" <<'PY'
import json,sys
try:
    d=json.load(open(sys.argv[1])); print(d["value"], d["ms_per_step"], "e2e", d["e2e"]["value"], d["step_mode"][:100], d.get("gradient_exchange",{}).get("note"), d.get("invalid"))
except Exception as e: print("ERR", e)
PY
done
tail -3 gpurun_out/r2j_bench2_peer.err
echo "total elapsed=$(( $(date +%s) - T0 ))s"
