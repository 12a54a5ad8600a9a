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
x = torch.randn(T, 768, device="cuda").bfloat16()
w = torch.randn(2304, 768, device="cuda").bfloat16()
for tn, cl in ((256, 1), (256, 2), (128, 1)):
    a = tm(lambda: ops.gemm(x, w, tile_n=tn, cluster=cl))
    b = tm(lambda: ops.gemm(x, w, tile_n=tn, cluster=cl, _debug_flags=1 << 29))
    c = tm(lambda: ops.gemm(x, w, tile_n=tn, cluster=cl, _debug_flags=1 << 30))
    o32 = torch.empty(T, 2304, device="cuda")
    d = tm(lambda: ops.gemm(x, w, tile_n=tn, cluster=cl, out=o32))
    print("tn%d/c%d: full %.1f | no-store %.1f | no-epilogue %.1f | fp32-out %.1f us" % (tn, cl, a, b, c, d), flush=True)
# same but small output (M=1024): tail effects
x = torch.randn(1024, 768, device="cuda").bfloat16()
for tn, cl in ((256, 1),):
    a = tm(lambda: ops.gemm(x, w, tile_n=tn, cluster=cl))
    b = tm(lambda: ops.gemm(x, w, tile_n=tn, cluster=cl, _debug_flags=1 << 29))
    c = tm(lambda: ops.gemm(x, w, tile_n=tn, cluster=cl, _debug_flags=1 << 30))
    print("M=1024 tn%d/c%d: full %.1f | no-store %.1f | no-epilogue %.1f us" % (tn, cl, a, b, c), flush=True)
