"""Thin Python wrappers over the C ABI (one function per ub200_* entry point).

These allocate outputs with torch (the ABI never allocates), pass raw pointers and launch on
torch's current CUDA stream.  No torch math happens here.
"""
import ctypes as C

import torch

from . import _lib
from ._lib import (EPI_ACCUM, EPI_ATOMIC, EPI_BIAS, EPI_COLSUM, EPI_DGELU, EPI_DROPOUT, EPI_GELU,
                   EPI_OUT_F32, EPI_RESIDUAL)


def gemm(a, b, *, a_major=0, b_major=0, bias=None, residual=None, aux=None, out=None,
         gelu=False, dgelu=False, accumulate=False, out_fp32=False, colsum=None,
         dropout_p=0.0, rng_seed=0, rng_stream=0, tile_n=0, max_ctas=0, cluster=0, k_splits=0,
         n_valid=0, rng_offset_dev=None, tanh=False, ln=None, _debug_flags=0):
    """D = epilogue(A . B^T) on the tcgen05 GEMM core.  Returns `out` (and pre-activation if gelu).

    a: [M,K] (a_major=0) or [K,M] (a_major=1);  b: [N,K] (b_major=0) or [K,N] (b_major=1).
    ln=(gamma, beta): fused residual + LayerNorm epilogue — returns (s, LayerNorm(s)); needs bias and
    residual, N = 768 or 1024.
    k_splits > 1 (or -1 = fill the SMs): split-K into a zero-initialised fp32 `out` through atomics.
    n_valid: b holds only n_valid of the N (= out.size(1)) output features; the rest get acc = 0.
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
    if n_valid:
        assert n_valid == N, "n_valid is the number of output features b really holds"
        N = out.size(1) if out is not None else (N + 7) // 8 * 8
    splitk = k_splits > 1 or k_splits == -1
    if out is None:
        if splitk:
            out = torch.zeros(M, N, device=a.device, dtype=torch.float32)
        else:
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
    if tanh:
        epi |= _lib.EPI_TANH
    ln_out = None
    if ln is not None:
        epi |= _lib.EPI_LN
        ln_out = torch.empty(M, N, device=a.device, dtype=a.dtype)
    if dgelu:
        epi |= EPI_DGELU
    if accumulate:
        epi |= EPI_ACCUM
    if out.dtype == torch.float32:
        epi |= EPI_OUT_F32
    if colsum is not None:
        epi |= EPI_COLSUM
    if splitk:
        assert out.dtype == torch.float32 and epi == EPI_OUT_F32, "split-K: fp32 out, no other epilogue"
        epi |= EPI_ATOMIC
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
        tile_n=int(tile_n), max_ctas=int(max_ctas), cluster=int(cluster), k_splits=int(k_splits),
        n_valid=int(n_valid), rng_offset_dev=rng_offset_dev,
        ln_gamma=_lib.ptr(ln[0]) if ln is not None else None,
        ln_beta=_lib.ptr(ln[1]) if ln is not None else None, ln_out=_lib.ptr(ln_out),
        ldln=ln_out.stride(0) if ln_out is not None else 0)
    _lib.check(lib.ub200_gemm(C.byref(args), _lib.current_stream()))
    if ln is not None:
        return out, ln_out
    return (out, out2) if gelu else out


def attn_fwd(qkv, cu_seqlens, max_seqlen, num_heads, dropout_p=0.0, rng_seed=0, rng_stream=0,
             rng_offset_dev=None):
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
                      rng_seed=int(rng_seed), rng_stream=int(rng_stream), rng_offset_dev=rng_offset_dev)
    _lib.check(lib.ub200_attn_fwd(C.byref(a), _lib.current_stream()))
    return ctx, lse


def attn_bwd(qkv, ctx, lse, dctx, cu_seqlens, max_seqlen, num_heads, dropout_p=0.0, rng_seed=0,
             rng_stream=0, dbias=None, rng_offset_dev=None):
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
                      workspace=ws.data_ptr() if ws_bytes else None, dbias=_lib.ptr(dbias),
                      rng_offset_dev=rng_offset_dev)
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
                  row_kind=None, kind=0, dropout_on_dy=False, dx=None, dgamma=None, dbeta=None, dbias=None,
                  zero_inactive=False, rng_offset_dev=None, split=False):
    """Returns dx, dx_drop (or None), dgamma, dbeta, dbias (fp32).  split=True: row kernel + column
    kernel (plain case only)."""
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
                       row_kind=_lib.ptr(row_kind), kind=int(kind),
                       dropout_on_dy=(1 if dropout_on_dy else 0) | (2 if zero_inactive else 0),
                       rng_offset_dev=rng_offset_dev)
    if split:
        ws = torch.empty(rows, 2, device=x.device, dtype=torch.float32)
        a.stats_ws = ws.data_ptr()
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


def cvt_from_f32(src, dtype, out=None, accumulate=False):
    """16-bit copy of an fp32 tensor (one launch)."""
    lib = _lib.load()
    src = src.contiguous()
    if out is None:
        out = torch.empty(src.shape, device=src.device, dtype=dtype)
    if src.numel():
        _lib.check(lib.ub200_cvt_from_f32(src.data_ptr(), out.data_ptr(), src.numel(),
                                          1 if accumulate else 0, _lib.dtype_code(dtype),
                                          _lib.current_stream()))
    return out


def ce_fwd(logits, targets, vocab):
    """loss [n] fp32, lse [n] fp32 of softmax cross-entropy over logits[:, :vocab] (16-bit, row
    pitch a multiple of 8)."""
    lib = _lib.load()
    n = logits.size(0)
    loss = torch.empty(n, device=logits.device, dtype=torch.float32)
    lse = torch.empty(n, device=logits.device, dtype=torch.float32)
    assert targets.dtype == torch.int64 and targets.is_contiguous() and logits.stride(1) == 1
    _lib.check(lib.ub200_ce_fwd(logits.data_ptr(), logits.stride(0), targets.data_ptr(), loss.data_ptr(),
                                lse.data_ptr(), n, vocab, _lib.dtype_code(logits.dtype),
                                _lib.current_stream()))
    return loss, lse


def ce_bwd_(logits, targets, lse, dloss, vocab):
    """In place: logits[:, c] <- (softmax - onehot) * dloss for c < vocab, 0 for the padding columns."""
    lib = _lib.load()
    n, ncols = logits.shape
    assert dloss.dtype == torch.float32 and dloss.is_contiguous()
    _lib.check(lib.ub200_ce_bwd(logits.data_ptr(), logits.data_ptr(), logits.stride(0), targets.data_ptr(),
                                lse.data_ptr(), dloss.data_ptr(), n, vocab, ncols,
                                _lib.dtype_code(logits.dtype), _lib.current_stream()))
    return logits


def dgelu_mul(dy, pre):
    lib = _lib.load()
    out = torch.empty_like(dy)
    assert dy.is_contiguous() and pre.is_contiguous() and dy.numel() % 8 == 0
    _lib.check(lib.ub200_dgelu_mul(dy.data_ptr(), pre.data_ptr(), out.data_ptr(), dy.numel(),
                                   _lib.dtype_code(dy.dtype), _lib.current_stream()))
    return out


def dtanh_mul(dy, y):
    """dy * (1 - y^2): backward of y = tanh(.) (BertPooler)."""
    lib = _lib.load()
    out = torch.empty_like(dy)
    assert dy.is_contiguous() and y.is_contiguous() and dy.numel() % 8 == 0
    _lib.check(lib.ub200_dtanh_mul(dy.data_ptr(), y.data_ptr(), out.data_ptr(), dy.numel(),
                                   _lib.dtype_code(dy.dtype), _lib.current_stream()))
    return out


def gather_rows(src, index, rows=None):
    """dst[r] = src[index[r]] if index[r] >= 0 else 0 (int32 index; bit-exact row mover)."""
    lib = _lib.load()
    src = src.contiguous()
    rows = index.numel() if rows is None else rows
    H = src.size(-1)
    dst = torch.empty(rows, H, device=src.device, dtype=src.dtype)
    if rows:
        _lib.check(lib.ub200_gather_rows(src.data_ptr(), dst.data_ptr(), index.data_ptr(), rows,
                                         H * src.element_size(), _lib.current_stream()))
    return dst
