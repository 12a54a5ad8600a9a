"""Thin Python wrappers over the C ABI (one function per ub200_* entry point).

These allocate outputs with torch (the ABI never allocates), pass raw pointers and launch on
torch's current CUDA stream.  No torch math happens here.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_ACCUM, EPI_BIAS, EPI_COLSUM, EPI_DGELU, EPI_DROPOUT, EPI_GELU,
                   EPI_OUT_F32, EPI_RESIDUAL)


def gemm(a, b, *, a_major=0, b_major=0, bias=None, residual=None, aux=None, out=None,
         gelu=False, dgelu=False, accumulate=False, out_fp32=False, colsum=None,
         dropout_p=0.0, rng_seed=0, rng_stream=0, tile_n=0, max_ctas=0, cluster=0, _debug_flags=0):
    """D = epilogue(A . B^T) on the tcgen05 GEMM core.  Returns `out` (and pre-activation if gelu).

    a: [M,K] (a_major=0) or [K,M] (a_major=1);  b: [N,K] (b_major=0) or [K,N] (b_major=1).
    """
    lib = _lib.load()
    assert a.is_cuda and b.is_cuda and a.dtype == b.dtype
    assert a.stride(-1) == 1 and b.stride(-1) == 1
    if a_major == 0:
        M, K = a.shape
    else:
        K, M = a.shape
    if b_major == 0:
        N, Kb = b.shape
    else:
        Kb, N = b.shape
    assert K == Kb, "contraction mismatch %d vs %d" % (K, Kb)
    if out is None:
        out = torch.empty(M, N, device=a.device, dtype=torch.float32 if out_fp32 else a.dtype)
    out2 = torch.empty(M, N, device=a.device, dtype=a.dtype) if gelu else None
    epi = 0
    if bias is not None:
        epi |= EPI_BIAS
    if dropout_p > 0:
        epi |= EPI_DROPOUT
    if residual is not None:
        epi |= EPI_RESIDUAL
    if gelu:
        epi |= EPI_GELU
    if dgelu:
        epi |= EPI_DGELU
    if accumulate:
        epi |= EPI_ACCUM
    if out.dtype == torch.float32:
        epi |= EPI_OUT_F32
    if colsum is not None:
        epi |= EPI_COLSUM
    epi |= _debug_flags
    args = _lib.GemmArgs(
        a=a.data_ptr(), b=b.data_ptr(), lda=a.stride(0), ldb=b.stride(0),
        a_major=a_major, b_major=b_major, M=M, N=N, K=K,
        dtype=_lib.dtype_code(a.dtype), epilogue=epi,
        bias=_lib.ptr(bias), residual=_lib.ptr(residual), aux=_lib.ptr(aux),
        out=out.data_ptr(), out2=_lib.ptr(out2), colsum=_lib.ptr(colsum),
        ldr=residual.stride(0) if residual is not None else 0,
        ldaux=aux.stride(0) if aux is not None else 0,
        ldo=out.stride(0),
        dropout_p=float(dropout_p), rng_seed=int(rng_seed), rng_stream=int(rng_stream),
        tile_n=int(tile_n), max_ctas=int(max_ctas), cluster=int(cluster))
    _lib.check(lib.ub200_gemm(C.byref(args), _lib.current_stream()))
    return (out, out2) if gelu else out


def attn_fwd(qkv, cu_seqlens, max_seqlen, num_heads, dropout_p=0.0, rng_seed=0, rng_stream=0):
    """ctx [T, H], lse [heads, T] = fused varlen attention over packed qkv [T, 3H]."""
    lib = _lib.load()
    T, H3 = qkv.shape
    H = H3 // 3
    ctx = torch.empty(T, H, device=qkv.device, dtype=qkv.dtype)
    lse = torch.empty(num_heads, T, device=qkv.device, dtype=torch.float32)
    a = _lib.AttnArgs(qkv=qkv.data_ptr(), ctx=ctx.data_ptr(), lse=lse.data_ptr(),
                      cu_seqlens=cu_seqlens.data_ptr(), batch=cu_seqlens.numel() - 1,
                      total_tokens=T, max_seqlen=max_seqlen, hidden=H, num_heads=num_heads,
                      dtype=_lib.dtype_code(qkv.dtype), dropout_p=float(dropout_p),
                      rng_seed=int(rng_seed), rng_stream=int(rng_stream))
    _lib.check(lib.ub200_attn_fwd(C.byref(a), _lib.current_stream()))
    return ctx, lse


def attn_bwd(qkv, ctx, lse, dctx, cu_seqlens, max_seqlen, num_heads, dropout_p=0.0, rng_seed=0,
             rng_stream=0):
    lib = _lib.load()
    T, H3 = qkv.shape
    H = H3 // 3
    dqkv = torch.empty_like(qkv)
    ws_bytes = lib.ub200_attn_bwd_workspace_bytes(T, H, max_seqlen)
    ws = torch.empty(max(ws_bytes, 1), device=qkv.device, dtype=torch.uint8)
    a = _lib.AttnArgs(qkv=qkv.data_ptr(), ctx=ctx.data_ptr(), lse=lse.data_ptr(),
                      cu_seqlens=cu_seqlens.data_ptr(), batch=cu_seqlens.numel() - 1,
                      total_tokens=T, max_seqlen=max_seqlen, hidden=H, num_heads=num_heads,
                      dtype=_lib.dtype_code(qkv.dtype), dropout_p=float(dropout_p),
                      rng_seed=int(rng_seed), rng_stream=int(rng_stream),
                      dctx=dctx.data_ptr(), dqkv=dqkv.data_ptr(),
                      workspace=ws.data_ptr() if ws_bytes else None)
    _lib.check(lib.ub200_attn_bwd(C.byref(a), _lib.current_stream()))
    return dqkv


def layernorm_fwd(x, gamma, beta):
    lib = _lib.load()
    rows, H = x.shape
    y = torch.empty_like(x)
    _lib.check(lib.ub200_layernorm_fwd(x.data_ptr(), gamma.data_ptr(), beta.data_ptr(), y.data_ptr(),
                                       rows, H, _lib.dtype_code(x.dtype), _lib.current_stream()))
    return y


def layernorm_bwd(dy, x, gamma, dropout_p=0.0, rng_seed=0, rng_stream=0, want_dbias=True,
                  row_kind=None, kind=0, dropout_on_dy=False, dx=None, dgamma=None, dbeta=None, dbias=None):
    """Returns dx, dx_drop (or None), dgamma, dbeta, dbias (fp32)."""
    lib = _lib.load()
    rows, H = x.shape
    if dx is None:
        dx = torch.empty_like(x) if row_kind is None else torch.zeros_like(x)
    dx_drop = torch.empty_like(x) if (dropout_p > 0 and not dropout_on_dy) else None
    if dgamma is None:
        dgamma = torch.zeros(H, device=x.device, dtype=torch.float32)
    if dbeta is None:
        dbeta = torch.zeros(H, device=x.device, dtype=torch.float32)
    if dbias is None and want_dbias:
        dbias = torch.zeros(H, device=x.device, dtype=torch.float32)
    a = _lib.LnBwdArgs(dy=dy.data_ptr(), x=x.data_ptr(), gamma=gamma.data_ptr(), dx=dx.data_ptr(),
                       dx_drop=_lib.ptr(dx_drop), dgamma=dgamma.data_ptr(), dbeta=dbeta.data_ptr(),
                       dbias=_lib.ptr(dbias), rows=rows, hidden=H, dtype=_lib.dtype_code(x.dtype),
                       dropout_p=float(dropout_p), rng_seed=int(rng_seed), rng_stream=int(rng_stream),
                       row_kind=_lib.ptr(row_kind), kind=int(kind), dropout_on_dy=1 if dropout_on_dy else 0)
    _lib.check(lib.ub200_layernorm_bwd(C.byref(a), _lib.current_stream()))
    return dx, dx_drop, dgamma, dbeta, dbias


def colsum(x, out=None):
    lib = _lib.load()
    rows, N = x.shape
    if out is None:
        out = torch.zeros(N, device=x.device, dtype=torch.float32)
    _lib.check(lib.ub200_colsum(x.data_ptr(), out.data_ptr(), rows, N, x.stride(0),
                                _lib.dtype_code(x.dtype), _lib.current_stream()))
    return out
