"""GPU: row-wise kernels (LayerNorm fwd/bwd with fused reductions, row gather, column sum)."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("split", [False, True])
@pytest.mark.parametrize("dtype,tol", [(torch.bfloat16, 2e-2), (torch.float16, 3e-3)])
@pytest.mark.parametrize("rows,H", [(1, 768), (3451, 768), (517, 1024), (33, 128)])
def test_layernorm_fwd_bwd(dtype, tol, rows, H, split):
    from uniter_b200 import ops
    torch.manual_seed(rows + H)
    x = (torch.randn(rows, H, device="cuda") * 2 + 0.3).to(dtype)
    g = (1 + 0.1 * torch.randn(H, device="cuda")).to(dtype)
    b = (0.02 * torch.randn(H, device="cuda")).to(dtype)
    dy = torch.randn(rows, H, device="cuda").to(dtype)
    x32 = x.float().requires_grad_(True)
    g32 = g.float().requires_grad_(True)
    b32 = b.float().requires_grad_(True)
    ref = torch.nn.functional.layer_norm(x32, (H,), g32, b32, eps=1e-12)
    ref.backward(dy.float())
    y = ops.layernorm_fwd(x, g, b)
    assert (y.float() - ref).abs().max().item() <= tol * max(1, ref.abs().max().item())
    dx, dxd, dg, db, dbias = ops.layernorm_bwd(dy, x, g, split=split)   # split: row + column kernels
    assert dxd is None
    assert (dx.float() - x32.grad).abs().max().item() <= tol * max(1, x32.grad.abs().max().item())
    assert (dg - g32.grad).abs().max().item() <= 2e-3 * max(1, g32.grad.abs().max().item())
    assert (db - b32.grad).abs().max().item() <= 2e-3 * max(1, b32.grad.abs().max().item())
    # dbias = column sum of dx as rounded to 16 bit
    assert (dbias - dx.float().sum(0)).abs().max().item() <= 1e-3 * max(1, dbias.abs().max().item())


@pytest.mark.parametrize("split", [False, True])
def test_layernorm_bwd_dropout_mask_matches_gemm_epilogue(split):
    """The backward must regenerate exactly the mask the forward GEMM epilogue applied."""
    from uniter_b200 import ops
    torch.manual_seed(0)
    rows, H = 640, 768
    x = torch.randn(rows, H, device="cuda").bfloat16()
    eye = torch.eye(H, device="cuda").bfloat16()
    fwd = ops.gemm(x, eye, dropout_p=0.1, rng_seed=77, rng_stream=9)   # x * mask / keep
    mask = fwd != 0
    dy = torch.randn(rows, H, device="cuda").bfloat16()
    g = torch.ones(H, device="cuda").bfloat16()
    dx, dxd, _, _, dbias = ops.layernorm_bwd(dy, x, g, dropout_p=0.1, rng_seed=77, rng_stream=9, split=split)
    nz = x != 0
    assert torch.equal((dxd != 0) | ~nz | (dx == 0), mask | ~nz | (dx == 0))
    kept = dxd != 0
    ratio = (dxd.float()[kept] / dx.float()[kept])
    assert (ratio - 1 / 0.9).abs().max().item() < 0.02
    assert (dbias - dxd.float().sum(0)).abs().max().item() < 1e-2


def test_gather_rows_bit_exact():
    import ctypes as C
    from uniter_b200 import _lib
    from uniter_b200.model import _bind
    lib = _bind()
    torch.manual_seed(3)
    src = torch.randn(1000, 768, device="cuda").half()
    idx = torch.randint(-1, 1000, (3000,), device="cuda", dtype=torch.int32)
    dst = torch.empty(3000, 768, device="cuda", dtype=torch.float16)
    _lib.check(lib.ub200_gather_rows(src.data_ptr(), dst.data_ptr(), idx.data_ptr(), 3000, 768 * 2,
                                     _lib.current_stream()))
    ref = torch.where((idx >= 0)[:, None], src[idx.clamp(min=0).long()], torch.zeros_like(dst))
    assert torch.equal(dst, ref)


@pytest.mark.parametrize("rows,N", [(3451, 2304), (5, 64), (1000, 3072)])
def test_colsum(rows, N):
    from uniter_b200 import ops
    x = torch.randn(rows, N, device="cuda").bfloat16()
    out = ops.colsum(x)
    ref = x.float().sum(0)
    assert (out - ref).abs().max().item() <= 1e-3 * max(1.0, ref.abs().max().item())
