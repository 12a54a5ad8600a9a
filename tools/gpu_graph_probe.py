"""Probe: can the current fwd+bwd step be captured in a CUDA graph as is, and what does replay cost?
(stale dropout masks / exact-T only — this is a feasibility + timing probe, not the product path)"""
import json
import sys
import time

import torch

sys.path.insert(0, ".")
from uniter_b200.heads import UniterForMLM  # noqa: E402
from uniter_b200.model import UniterConfig, register_lengths  # noqa: E402
from uniter_b200.synth import synth_batch  # noqa: E402

dev = torch.device("cuda", 0)
torch.manual_seed(0)
cfg = UniterConfig(28996, hidden_size=768, num_hidden_layers=12, num_attention_heads=12,
                   intermediate_size=3072, max_position_embeddings=512)
model = UniterForMLM(cfg, 2048).to(device=dev, dtype=torch.bfloat16).train()
hb = synth_batch(64, 12, 28, 26, 46, 1234, mlm_prob=0.15)
batch = {k: v.to(dev) for k, v in hb.items() if torch.is_tensor(v)}
register_lengths(batch["attn_masks"], [a + b for a, b in zip(hb["txt_lens"], hb["num_bbs"])], prefix=True)


def step():
    model.zero_grad(set_to_none=True)
    loss = model(batch).mean()
    loss.backward()
    return loss


def timeit(fn, n):
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    t0 = time.perf_counter()
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    host = (time.perf_counter() - t0) / n * 1e3
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n, host


s = torch.cuda.Stream()
s.wait_stream(torch.cuda.current_stream())
with torch.cuda.stream(s):
    for _ in range(5):
        step()
torch.cuda.current_stream().wait_stream(s)
res = {}
res["eager_ms"], res["eager_host_ms"] = timeit(step, 20)
g = torch.cuda.CUDAGraph()
try:
    model.zero_grad(set_to_none=True)
    with torch.cuda.graph(g):
        loss = step()
    res["captured"] = True
    for _ in range(3):
        g.replay()
    res["graph_ms"], res["graph_host_ms"] = timeit(g.replay, 50)
    res["loss"] = float(loss)
except Exception as e:  # noqa: BLE001
    res["captured"] = False
    res["error"] = "%s: %s" % (type(e).__name__, e)
print(json.dumps(res))
