"""Host-side profile of one C2 training step (cProfile over the enqueue path; GPU runs async).
Usage (GPU box): python tools/gpu_host_prof.py > gpurun_out/host_prof.txt"""
import cProfile, pstats, io, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from bench import BASE, C2
from uniter_b200.model import UniterConfig, register_lengths
from uniter_b200.heads import UniterForMLM
from uniter_b200.synth import synth_batch

dev = torch.device("cuda", 0)
torch.manual_seed(0)
cfg = UniterConfig(BASE["vocab"], hidden_size=BASE["H"], num_hidden_layers=12,
                   num_attention_heads=BASE["heads"], intermediate_size=BASE["I"],
                   max_position_embeddings=BASE["max_pos"])
model = UniterForMLM(cfg, BASE["img_dim"]).to(device=dev, dtype=torch.bfloat16).train()
b = synth_batch(C2["B"], C2["tl"][0], C2["tl"][1], C2["nbb"][0], C2["nbb"][1], C2["seed"],
                mlm_prob=C2["mlm_prob"])
d = {k: v.to(dev) for k, v in b.items() if torch.is_tensor(v)}
register_lengths(d["attn_masks"], [x + y for x, y in zip(b["txt_lens"], b["num_bbs"])], prefix=True)


def step():
    model.zero_grad(set_to_none=True)
    loss = model(d).mean()
    loss.backward()


for _ in range(5):
    step()
torch.cuda.synchronize()
# phase timing
N = 30
tz = tf = tb = 0.0
for _ in range(N):
    t0 = time.perf_counter(); model.zero_grad(set_to_none=True)
    t1 = time.perf_counter(); loss = model(d).mean()
    t2 = time.perf_counter(); loss.backward()
    t3 = time.perf_counter()
    tz += t1 - t0; tf += t2 - t1; tb += t3 - t2
    if _ % 4 == 3:
        torch.cuda.synchronize()
print("host ms/step: zero_grad %.3f  forward %.3f  backward %.3f" % (tz / N * 1e3, tf / N * 1e3, tb / N * 1e3))
torch.cuda.synchronize()
pr = cProfile.Profile()
pr.enable()
for i in range(20):
    step()
    if i % 4 == 3:
        torch.cuda.synchronize()
pr.disable()
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(45)
print(s.getvalue())
s = io.StringIO()
pstats.Stats(pr, stream=s).sort_stats("cumulative").print_stats(45)
print(s.getvalue())
