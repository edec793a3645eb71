"""LayerNorm kernels against torch autograd (fp32), through the C ABI."""
import pytest
import torch

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("rows,cols", [(1, 1936), (7, 1936), (257, 1936), (1000, 2048), (3000, 512), (5, 8)])
def test_layernorm_bwd_fused_matches_autograd(cuda_lib, rows, cols):
    """One-pass backward (dx, bf16 copy, dw, db and the column sums of dx) vs torch, and vs the two-kernel form."""
    from nlvsgg_b200 import ops
    g = torch.Generator().manual_seed(rows * 31 + cols)
    x = torch.randn(rows, cols, generator=g).cuda().requires_grad_(True)
    w = (1 + 0.1 * torch.randn(cols, generator=g)).cuda().requires_grad_(True)
    b = torch.randn(cols, generator=g).cuda().requires_grad_(True)
    dy = torch.randn(rows, cols, generator=g).cuda()
    torch.nn.functional.layer_norm(x, (cols,), w, b, 1e-5).backward(dy)
    _, _, mean, rstd = ops.layernorm_fwd(x.detach(), w.detach(), b.detach())
    dx, dx2, dw, db, dprev = ops.layernorm_bwd_fused(dy, x.detach(), mean, rstd, w.detach(), dx2_dtype=torch.bfloat16)
    scale = x.grad.abs().max().item()
    assert (dx - x.grad).abs().max().item() <= 2e-5 * scale + 1e-6
    assert torch.equal(dx2, dx.bfloat16())
    assert (dw - w.grad).abs().max().item() <= 1e-4 * w.grad.abs().max().item() + 1e-5
    assert (db - b.grad).abs().max().item() <= 1e-4 * b.grad.abs().max().item() + 1e-5
    want = dx.double().sum(0)
    assert (dprev.double() - want).abs().max().item() <= 1e-4 * want.abs().max().item() + 1e-5
    dx_b, _, dw_b, db_b = ops.layernorm_bwd(dy, x.detach(), mean, rstd, w.detach())
    assert (dx - dx_b).abs().max().item() <= 2e-5 * scale + 1e-6 and (dw - dw_b).abs().max().item() <= 1e-4 * dw_b.abs().max().item() + 1e-5


def test_layernorm_bwd_fused_dropout_operand(cuda_lib):
    from nlvsgg_b200 import _C, ops
    g = torch.Generator().manual_seed(4)
    rows, cols, p = 301, 1936, 0.1
    x, dy = torch.randn(rows, cols, generator=g).cuda(), torch.randn(rows, cols, generator=g).cuda()
    w, b = torch.randn(cols, generator=g).cuda(), torch.randn(cols, generator=g).cuda()
    _, _, mean, rstd = ops.layernorm_fwd(x, w, b)
    d = _C.Dropout.make(p, 77, 3)
    dx, dx2, _, _, dprev = ops.layernorm_bwd_fused(dy, x, mean, rstd, w, dx2_dtype=torch.float32, drop=d)
    dx_plain, _, _, _, _ = ops.layernorm_bwd_fused(dy, x, mean, rstd, w)
    assert torch.equal(dx, dx_plain)                                    # the residual branch stays unmasked
    mask = ops.dropout_mask(rows, cols, d).float()
    assert (dx2 - dx * mask / (1 - p)).abs().max().item() < 1e-6
    want = dx2.double().sum(0)
    assert (dprev.double() - want).abs().max().item() <= 1e-4 * want.abs().max().item() + 1e-5
