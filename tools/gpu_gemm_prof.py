import sys, torch
sys.path.insert(0, ".")
from uniter_b200 import ops
T = 3451
x = torch.randn(T, 768, device="cuda").bfloat16()
w = torch.randn(2304, 768, device="cuda").bfloat16()
bias = torch.randn(2304, device="cuda").bfloat16()
for _ in range(3):
    ops.gemm(x, w, bias=bias, tile_n=256, cluster=1)
torch.cuda.synchronize()
