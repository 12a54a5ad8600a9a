import sys, torch
sys.path.insert(0, ".")
from uniter_b200 import ops
T = 3451
x = torch.randn(T, 768, device="cuda").bfloat16()
w = torch.randn(2304, 768, device="cuda").bfloat16()
names = ["start", "setup_done", "first_tma", "producer_done", "first_full", "mma_issued_all", "epi_t0_begin",
         "epi_t0_end", "epi_t1_begin", "epi_t1_end", "w11_end", "cta_end"]
for tn, cl in ((256, 1),):
  for label, flags in (("normal", 0), ("skip-ldtm", 1 << 28), ("no-store", 1 << 29), ("no-epilogue", 1 << 30)):
    for _ in range(3):
        ops.gemm(x, w, tile_n=tn, cluster=cl, _debug_flags=flags)
    buf = torch.zeros(148 * 16 * 2 + 4096, device="cuda", dtype=torch.float32)
    # colsum pointer doubles as trace buffer; EPI_COLSUM stays off because we pass it via aux-less path
    from uniter_b200 import _lib
    import ctypes as C
    out = torch.empty(T, 2304, device="cuda", dtype=torch.bfloat16)
    args = _lib.GemmArgs(a=x.data_ptr(), b=w.data_ptr(), lda=768, ldb=768, a_major=0, b_major=0, M=T, N=2304, K=768,
                         dtype=_lib.BF16, epilogue=(1 << 27) | flags, out=out.data_ptr(), colsum=buf.data_ptr(), ldo=2304,
                         tile_n=tn, cluster=cl)
    _lib.check(_lib.load().ub200_gemm(C.byref(args), _lib.current_stream()))
    torch.cuda.synchronize()
    t = buf[:148 * 32].view(torch.int64).view(148, 16).cpu()
    print("==", label, "tn", tn, "cluster", cl)
    for b in (0, 94, 95, 147):
        base = t[b, 0].item()
        print("cta %3d:" % b, "  ".join("%s=%d" % (n, t[b, i].item() - base if t[b, i].item() else -1) for i, n in enumerate(names)))
