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
         dropout_p=0.0, rng_seed=0, rng_stream=0, tile_n=0, max_ctas=0):
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
        tile_n=int(tile_n), max_ctas=int(max_ctas))
    _lib.check(lib.ub200_gemm(C.byref(args), _lib.current_stream()))
    return (out, out2) if gelu else out
