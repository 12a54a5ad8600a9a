import sys, torch
sys.path.insert(0, ".")
from uniter_b200 import ops
def tm(fn, n=30):
    for _ in range(5): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n * 1e3
T = 3451
for K in (768, 3072, 12288):
    x = torch.randn(T, K, device="cuda").bfloat16()
    w = torch.randn(2304, K, device="cuda").bfloat16()
    bias = torch.randn(2304, device="cuda").bfloat16()
    for tn, cl in ((256, 1), (128, 1), (256, 2), (128, 2)):
        a = tm(lambda: ops.gemm(x, w, tile_n=tn, cluster=cl))
        b = tm(lambda: ops.gemm(x, w, tile_n=tn, cluster=cl, _debug_flags=1 << 30))
        c = tm(lambda: ops.gemm(x, w, bias=bias, tile_n=tn, cluster=cl))
        print("K=%d tn%d/c%d: full %.1f us | no-epilogue %.1f us | bias %.1f us  (%.0f TF full)" % (K, tn, cl, a, b, c, 2.0*T*2304*K/a/1e6), flush=True)
    us = tm(lambda: x @ w.t())
    print("K=%d cublas %.1f us %.0f TF" % (K, us, 2.0*T*2304*K/us/1e6), flush=True)
# empty kernel launch overhead reference
y = torch.zeros(8, device="cuda")
print("tiny torch kernel: %.1f us" % tm(lambda: y.add_(1)))
