"""GPU bring-up diagnostic for the tcgen05 GEMM core (run under gpurun).

Checks every operand-major combination / tile width / dtype against torch fp32 matmul and
prints the error structure so that a descriptor mistake can be diagnosed from one trip.
"""
import sys
import time

import torch

sys.path.insert(0, ".")
from uniter_b200 import _lib, ops  # noqa: E402


def describe(name, got, ref, tol):
    got = got.float()
    err = (got - ref).abs()
    mx = err.max().item()
    ok = mx <= tol * max(1.0, ref.abs().max().item())
    msg = "%-48s max_err=%.4e ref_max=%.3e %s" % (name, mx, ref.abs().max().item(),
                                                   "OK" if ok else "FAIL")
    if not ok:
        bad = (err > tol * max(1.0, ref.abs().max().item()))
        rows = bad.any(1).nonzero().flatten()
        cols = bad.any(0).nonzero().flatten()
        msg += "\n    bad frac=%.4f rows[%d]: %s cols[%d]: %s" % (
            bad.float().mean().item(), rows.numel(), rows[:12].tolist(), cols.numel(),
            cols[:12].tolist())
        msg += "\n    got[0,:8]=%s\n    ref[0,:8]=%s" % (got[0, :8].tolist(), ref[0, :8].tolist())
        nz = (got != 0).float().mean().item()
        msg += "\n    nonzero frac of got=%.4f  nan=%d" % (nz, int(torch.isnan(got).sum()))
    print(msg, flush=True)
    return ok


def main():
    lib = _lib.load()
    print("version", lib.ub200_version(), "device_check", lib.ub200_device_check(),
          torch.cuda.get_device_name(0), flush=True)
    torch.manual_seed(0)
    dev = "cuda"
    all_ok = True
    for dtype in (torch.bfloat16, torch.float16):
        tol = 2e-2 if dtype == torch.bfloat16 else 4e-3
        for (M, N, K) in ((128, 128, 64), (256, 256, 256), (333, 768, 768), (3451, 2304, 768),
                          (777, 768, 3072)):
            x = torch.randn(M, K, device=dev).to(dtype)
            w = (torch.randn(N, K, device=dev) * 0.05).to(dtype)
            ref = x.float() @ w.float().t()
            for tn in (64, 128, 256):
                # forward: both K-major
                out = ops.gemm(x, w, tile_n=tn)
                torch.cuda.synchronize()
                all_ok &= describe("KK  %s %dx%dx%d tn=%d" % (str(dtype)[6:], M, N, K, tn), out, ref, tol)
                # dgrad form: B MN-major  (w stored [K_contract, N_out]) -> D = x @ wt^T?  use wt=[K,N]
                wt = w.t().contiguous()  # [K, N]
                out = ops.gemm(x, wt, b_major=1, tile_n=tn)
                torch.cuda.synchronize()
                all_ok &= describe("K,MN %s %dx%dx%d tn=%d" % (str(dtype)[6:], M, N, K, tn), out, ref, tol)
                # wgrad form: both MN-major
                Mp = (M + 7) // 8 * 8  # TMA needs pitch % 8 == 0
                xt = torch.zeros(K, Mp, device=dev, dtype=dtype)[:, :M]
                xt.copy_(x.t())
                out = ops.gemm(xt, wt, a_major=1, b_major=1, tile_n=tn)
                torch.cuda.synchronize()
                all_ok &= describe("MN,MN %s %dx%dx%d tn=%d" % (str(dtype)[6:], M, N, K, tn), out, ref, tol)
        # ragged K (wgrad contraction over T)
        M, N, K = 768, 768, 3451
        xt = torch.randn(K, M, device=dev).to(dtype)
        wt = (torch.randn(K, N, device=dev) * 0.05).to(dtype)
        ref = xt.float().t() @ wt.float()
        out = ops.gemm(xt, wt, a_major=1, b_major=1, out_fp32=True)
        torch.cuda.synchronize()
        all_ok &= describe("MN,MN raggedK f32out %s" % str(dtype)[6:], out, ref, tol)

        # epilogues
        M, N, K = 515, 768, 768
        x = torch.randn(M, K, device=dev).to(dtype)
        w = (torch.randn(N, K, device=dev) * 0.05).to(dtype)
        bias = torch.randn(N, device=dev).to(dtype)
        res = torch.randn(M, N, device=dev).to(dtype)
        base = x.float() @ w.float().t()
        out = ops.gemm(x, w, bias=bias)
        all_ok &= describe("epi bias", out, base + bias.float(), tol)
        out = ops.gemm(x, w, bias=bias, residual=res)
        all_ok &= describe("epi bias+res", out, base + bias.float() + res.float(), tol)
        out, pre = ops.gemm(x, w, bias=bias, gelu=True)
        refpre = base + bias.float()
        all_ok &= describe("epi gelu(pre)", pre, refpre, tol)
        all_ok &= describe("epi gelu(act)", out, torch.nn.functional.gelu(refpre), tol)
        aux = torch.randn(M, N, device=dev).to(dtype)
        a32 = aux.float().requires_grad_(True)
        torch.nn.functional.gelu(a32).sum().backward()
        out = ops.gemm(x, w, aux=aux, dgelu=True)
        all_ok &= describe("epi dgelu", out, base * a32.grad, tol)
        cs = torch.zeros(N, device=dev)
        out = ops.gemm(x, w, bias=bias, colsum=cs)
        all_ok &= describe("epi colsum", cs[None], (base + bias.float()).sum(0)[None], tol)
        acc = torch.randn(M, N, device=dev)
        acc0 = acc.clone()
        ops.gemm(x, w, out=acc, accumulate=True)
        all_ok &= describe("epi accum f32", acc, base + acc0, tol)
        acc = torch.randn(M, N, device=dev).to(dtype)
        acc0 = acc.clone()
        ops.gemm(x, w, out=acc, accumulate=True)
        all_ok &= describe("epi accum 16b", acc, base + acc0.float(), tol)
        # dropout statistics + determinism
        out = ops.gemm(x, w, bias=bias, dropout_p=0.1, rng_seed=123, rng_stream=7)
        out_b = ops.gemm(x, w, bias=bias, dropout_p=0.1, rng_seed=123, rng_stream=7)
        out_c = ops.gemm(x, w, bias=bias, dropout_p=0.1, rng_seed=123, rng_stream=8)
        dropped = (out == 0).float().mean().item()
        same = torch.equal(out, out_b)
        diff = (out != out_c).float().mean().item()
        kept = out != 0
        ref = (base + bias.float()) / 0.9
        err = ((out.float() - ref).abs() * kept).max().item()
        print("dropout: dropped=%.4f (want ~0.1) deterministic=%s other-stream-diff=%.3f kept_err=%.3e"
              % (dropped, same, diff, err), flush=True)
        all_ok &= abs(dropped - 0.1) < 0.01 and same and diff > 0.1 and err < tol * 10

    # timing of the forward / dgrad / wgrad GEMM shapes at C2 (T=3451)
    def tm(fn, n=30):
        for _ in range(5):
            fn()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
        e0.record()
        for _ in range(n):
            fn()
        e1.record()
        torch.cuda.synchronize()
        return e0.elapsed_time(e1) / n * 1e3
    T = 3451
    for (M, N, K, am, bm, what) in ((T, 2304, 768, 0, 0, "qkv fwd"), (T, 768, 768, 0, 0, "attnout fwd"),
                                    (T, 3072, 768, 0, 0, "ffn1 fwd"), (T, 768, 3072, 0, 0, "ffn2 fwd"),
                                    (T, 3072, 768, 0, 1, "ffn2 dgrad"), (T, 768, 3072, 0, 1, "ffn1 dgrad"),
                                    (768, 3072, T, 1, 1, "ffn2 wgrad"), (3072, 768, T, 1, 1, "ffn1 wgrad"),
                                    (768, 768, T, 1, 1, "attnout wgrad"), (2304, 768, T, 1, 1, "qkv wgrad")):
        a = torch.randn((K, M) if am else (M, K), device=dev).to(torch.bfloat16)
        b = torch.randn((K, N) if bm else (N, K), device=dev).to(torch.bfloat16)
        res = []
        for tn, cl in ((0, 0), (128, 1), (192, 1), (256, 1), (128, 2), (256, 2)):
            us = tm(lambda: ops.gemm(a, b, a_major=am, b_major=bm, tile_n=tn, cluster=cl))
            res.append("tn%d/c%d %.1fus %.0fTF" % (tn, cl, us, 2.0 * M * N * K / us / 1e6))
        aa = a.t() if am else a
        bb = b if bm else b.t()
        us = tm(lambda: aa @ bb)
        print("time %-14s %dx%dx%d: %s | cublas %.1fus %.0fTF" % (what, M, N, K, "  ".join(res), us,
                                                              2.0 * M * N * K / us / 1e6), flush=True)
    # epilogue cost: ffn1 with GELU, ffn2 dgrad with dGELU + colsum
    x = torch.randn(T, 768, device=dev).bfloat16(); w = torch.randn(3072, 768, device=dev).bfloat16()
    bias = torch.randn(3072, device=dev).bfloat16()
    print("ffn1 gelu epilogue: %.1f us" % tm(lambda: ops.gemm(x, w, bias=bias, gelu=True)), flush=True)
    dy = torch.randn(T, 768, device=dev).bfloat16(); w2 = torch.randn(768, 3072, device=dev).bfloat16()
    pre = torch.randn(T, 3072, device=dev).bfloat16(); cs = torch.zeros(3072, device=dev)
    print("ffn2 dgrad dgelu+colsum epilogue: %.1f us" % tm(lambda: ops.gemm(dy, w2, b_major=1, aux=pre, dgelu=True, colsum=cs)), flush=True)
    print("ALL_OK" if all_ok else "SOME_FAILED", flush=True)


if __name__ == "__main__":
    t0 = time.time()
    main()
    print("elapsed %.1fs" % (time.time() - t0))
