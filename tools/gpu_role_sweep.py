"""Per-role GEMM sweep on the C2 shapes: every (N tile, cluster) the dispatcher accepts, with the
role's REAL epilogue, timed back to back with CUDA events (kernel-only, PDL on, L2-warm like the
steady-state step).  Prints one table per role and writes gpurun_out/role_sweep.json.

    python tools/gpu_role_sweep.py [T]           (GPU box; ~15 s)

Use it to re-fit pick_config()'s cost model after a GEMM change, and to see which roles are
latency-bound (time barely moves with the tile) vs throughput-bound."""
import json
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402

from uniter_b200 import ops  # noqa: E402

T = int(sys.argv[1]) if len(sys.argv) > 1 else 3451
H, I = 768, 3072
dt = torch.bfloat16
dev = "cuda"


def tm(fn, n=40):
    for _ in range(5):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3


def rnd(*shape):
    return (torch.randn(*shape, device=dev) * 0.05).to(dt)


x, ctx, a, f, pre = rnd(T, H), rnd(T, H), rnd(T, H), rnd(T, I), rnd(T, I)
dy, dpre, dqkv = rnd(T, H), rnd(T, I), rnd(T, 3 * H)
wqkv, wo, w1, w2 = rnd(3 * H, H), rnd(H, H), rnd(I, H), rnd(H, I)
bqkv, bo, b1, b2 = rnd(3 * H), rnd(H), rnd(I), rnd(H)
cs = torch.zeros(I, device=dev)

ROLES = {
    # name: (flops, callable(tile_n, cluster))
    "qkv_fwd  [T,768]x[2304,768]^T +bias": (2 * T * H * 3 * H, lambda tn, cl: ops.gemm(x, wqkv, bias=bqkv, tile_n=tn, cluster=cl)),
    "attnout_fwd +bias+dropout+residual": (2 * T * H * H, lambda tn, cl: ops.gemm(ctx, wo, bias=bo, residual=x, dropout_p=0.1, rng_seed=1, rng_stream=2, tile_n=tn, cluster=cl)),
    "ffn1_fwd +bias+gelu (2 outputs)": (2 * T * H * I, lambda tn, cl: ops.gemm(a, w1, bias=b1, gelu=True, tile_n=tn, cluster=cl)),
    "ffn2_fwd K=3072 +bias+dropout+residual": (2 * T * H * I, lambda tn, cl: ops.gemm(f, w2, bias=b2, residual=a, dropout_p=0.1, rng_seed=1, rng_stream=3, tile_n=tn, cluster=cl)),
    "ffn2_dgrad +dgelu+colsum": (2 * T * H * I, lambda tn, cl: ops.gemm(dy, w2, b_major=1, aux=pre, dgelu=True, colsum=cs, tile_n=tn, cluster=cl)),
    "ffn1_dgrad K=3072 +residual": (2 * T * H * I, lambda tn, cl: ops.gemm(dpre, w1, b_major=1, residual=dy, tile_n=tn, cluster=cl)),
    "attnout_dgrad": (2 * T * H * H, lambda tn, cl: ops.gemm(dy, wo, b_major=1, tile_n=tn, cluster=cl)),
    "qkv_dgrad K=2304 +residual": (2 * T * H * 3 * H, lambda tn, cl: ops.gemm(dqkv, wqkv, b_major=1, residual=dy, tile_n=tn, cluster=cl)),
    "wgrad dW1 [3072,768] K=T": (2 * T * H * I, lambda tn, cl: ops.gemm(dpre, a, a_major=1, b_major=1, tile_n=tn, cluster=cl)),
}
CONFIGS = [(0, 0), (64, 1), (128, 1), (192, 1), (256, 1), (128, 2), (256, 2)]

out = {}
for name, (flops, fn) in ROLES.items():
    row = {}
    for tn, cl in CONFIGS:
        try:
            us = tm(lambda: fn(tn, cl))
        except RuntimeError as e:
            row["%d/%d" % (tn, cl)] = None
            continue
        row["%d/%d" % (tn, cl)] = round(us, 2)
    out[name] = row
    best = min((v, k) for k, v in row.items() if v)
    print("%-42s %s   best %s: %.1f us = %.0f TF" % (
        name, "  ".join("%s=%s" % (k, "%.1f" % v if v else "n/a") for k, v in row.items()), best[1], best[0],
        flops / best[0] / 1e6), flush=True)
os.makedirs("gpurun_out", exist_ok=True)
with open("gpurun_out/role_sweep.json", "w") as fh:
    json.dump({"T": T, "us": out}, fh, indent=1)
