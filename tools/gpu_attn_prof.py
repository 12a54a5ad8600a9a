import sys, torch
sys.path.insert(0, ".")
from uniter_b200 import ops
from uniter_b200.synth import synth_batch
b = synth_batch(64, 12, 28, 26, 46, 1234)
lens = [a + c for a, c in zip(b["txt_lens"], b["num_bbs"])]
T, H, heads = sum(lens), 768, 12
qkv = torch.randn(T, 3 * H, device="cuda").bfloat16()
cu = torch.tensor([0] + list(torch.tensor(lens).cumsum(0)), device="cuda", dtype=torch.int32)
dctx = torch.randn(T, H, device="cuda").bfloat16()
for _ in range(3):
    ctx, lse = ops.attn_fwd(qkv, cu, max(lens), heads, dropout_p=0.1, rng_seed=1, rng_stream=2)
    dqkv = ops.attn_bwd(qkv, ctx, lse, dctx, cu, max(lens), heads, dropout_p=0.1, rng_seed=1, rng_stream=2)
torch.cuda.synchronize()
