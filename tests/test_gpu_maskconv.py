"""Fused first stage of the spatial-mask branch (csrc/maskconv.cu; lib/sttran.py:337-341) against torch and against the
im2col + GEMM + statistics + BatchNorm + pooling kernels it replaces, through the C ABI."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _masks(r, seed):
    """Box-shaped masks with fractional edges (what draw_union_boxes produces), plus a few dense random ones."""
    g = torch.Generator().manual_seed(seed)
    m = torch.zeros(r, 2, 27, 27)
    for i in range(r):
        for c in range(2):
            if i % 5 == 4:
                m[i, c] = torch.rand(27, 27, generator=g)
                continue
            x0, y0 = [int(v) for v in torch.randint(0, 20, (2,), generator=g)]
            x1, y1 = x0 + int(torch.randint(1, 27 - x0, (1,), generator=g)), y0 + int(torch.randint(1, 27 - y0, (1,), generator=g))
            m[i, c, y0:y1, x0:x1] = 1.0
            m[i, c, y0, x0:x1] = float(torch.rand(1, generator=g))
            m[i, c, y0:y1, x0] = float(torch.rand(1, generator=g))
    return m


def _videos(r, nv):
    """Pairs split over nv videos (one of them empty when nv > 2)."""
    cuts = sorted(int(v) for v in torch.randint(0, r + 1, (nv - 1,), generator=torch.Generator().manual_seed(r + nv)))
    bounds = [0] + cuts + [r]
    if nv > 2:
        bounds[2] = bounds[1]
    pair_video = torch.zeros(r, dtype=torch.int32)
    for v in range(nv):
        pair_video[bounds[v]:bounds[v + 1]] = v
    seg196 = torch.tensor([b * 196 for b in bounds], dtype=torch.int32)
    return pair_video.cuda(), seg196.cuda(), bounds


@pytest.mark.parametrize("r,nv", [(1, 1), (3, 1), (37, 3), (700, 5)])
def test_mask_conv1_fwd_matches_torch_and_gemm_path(cuda_lib, r, nv):
    from nlvsgg_b200 import ops
    g = torch.Generator().manual_seed(r)
    masks = _masks(r, r).cuda()
    w = (0.2 * torch.randn(128, 2, 7, 7, generator=g)).cuda()
    bias = (0.1 * torch.randn(128, generator=g)).cuda()
    pair_video, seg196, bounds = _videos(r, nv)
    rm, rv = torch.zeros(128).cuda(), torch.ones(128).cuda()
    out, mean, var = ops.mask_conv1_fwd(masks, w.view(128, 98), bias, pair_video, seg196, nv, 0.01, rm, rv)

    # torch on the bf16-rounded operands (what both kernel routes multiply), fp32 accumulation
    want = F.relu(F.conv2d(masks.bfloat16().float(), w.bfloat16().float(), bias, stride=2, padding=3))      # [r,128,14,14]
    want = want.permute(0, 2, 3, 1).reshape(r * 196, 128)
    assert (out.float() - want).abs().max().item() <= 1e-2 * want.abs().max().item()      # bf16 rounding of the stored map
    assert (out.float() - want.bfloat16().float()).abs().max().item() <= 2 ** -7 * want.abs().max().item()

    # the replaced route: im2col -> tcgen05 GEMM (+ bias + ReLU) -> nlv_bn_stats
    col = ops.im2col_mask(masks, torch.bfloat16)
    wop = torch.zeros(128, 104, device="cuda", dtype=torch.bfloat16)
    wop[:, :98] = w.view(128, 98).bfloat16()
    ref = torch.empty(r * 196, 128, device="cuda", dtype=torch.bfloat16)
    ops.gemm(col, wop, ref, bias=bias, relu=True)
    diff = (out.float() - ref.float()).abs()
    assert diff.max().item() <= 2 ** -7 * ref.float().abs().max().item()                  # at most one bf16 ulp (summation order)
    assert (diff > 0).float().mean().item() < 0.02
    rm2, rv2 = torch.zeros(128).cuda(), torch.ones(128).cuda()
    mean2, var2 = ops.bn_stats(out, seg196, nv, 128, 0.01, rm2, rv2)                       # statistics of the SAME stored map
    assert (mean - mean2).abs().max().item() <= 1e-5 * mean2.abs().max().item() + 1e-7
    assert (var - var2).abs().max().item() <= 1e-4 * var2.abs().max().item() + 1e-7
    assert (rm - rm2).abs().max().item() <= 1e-6 and (rv - rv2).abs().max().item() <= 1e-6
    for v in range(nv):                                                                    # and against torch, per video
        a, b = bounds[v] * 196, bounds[v + 1] * 196
        if b > a:
            x = out[a:b].double()
            assert (mean[v].double() - x.mean(0)).abs().max().item() <= 1e-5
            assert (var[v].double() - x.var(0, unbiased=False)).abs().max().item() <= 1e-5 * max(1.0, x.var(0).max().item())


@pytest.mark.parametrize("r,nv", [(1, 1), (37, 3), (300, 2)])
def test_bn_apply_maxpool_matches_the_two_kernels(cuda_lib, r, nv):
    from nlvsgg_b200 import ops
    g = torch.Generator().manual_seed(r + 1)
    x = torch.relu(torch.randn(r * 196, 128, generator=g)).bfloat16().cuda()
    x[::7] = x[1::7][: x[::7].shape[0]]                                                   # ties inside pooling windows
    pair_video, seg196, _ = _videos(r, nv)
    mean = torch.randn(nv, 128, generator=g).cuda() * 0.3
    var = (torch.rand(nv, 128, generator=g) + 0.05).cuda()
    w = torch.randn(128, generator=g).cuda()                                              # negative scales flip the ordering
    b = torch.randn(128, generator=g).cuda()
    y, arg = ops.bn_apply_maxpool(x, pair_video, mean, var, w, b, r)
    bn, _ = ops.bn_apply(x, pair_video, mean, var, w, b, False, out_dtype=torch.bfloat16, row_div=196)
    y2, arg2 = ops.maxpool_fwd(bn, r, 128, torch.bfloat16)
    assert torch.equal(y, y2)
    # the tap is an argmax of the window: it may differ from the two-kernel one only between outputs that tie after rounding
    rows = torch.arange(r * 49, device="cuda").view(-1, 1)
    pairs, cell, tap = rows // 49, rows % 49, arg.long()
    src = pairs * 196 + (2 * (cell // 7) - 1 + tap // 3) * 14 + (2 * (cell % 7) - 1 + tap % 3)
    assert torch.equal(bn[src, torch.arange(128, device="cuda").view(1, -1)], y)
    assert (arg != arg2).float().mean().item() < 0.05


@pytest.mark.parametrize("r", [1, 5, 333])
def test_mask_conv1_dw_matches_torch_and_gemm_path(cuda_lib, r):
    from nlvsgg_b200 import ops
    from nlvsgg_b200._C import MAJOR_MN
    g = torch.Generator().manual_seed(r + 2)
    masks = _masks(r, r + 9).cuda()
    dy = (torch.randn(r * 196, 128, generator=g) * (torch.rand(r * 196, 128, generator=g) > 0.4)).bfloat16().cuda()
    dw = ops.mask_conv1_dw(dy, masks)
    # torch: gradient of conv2d w.r.t. the weight on the bf16-rounded operands
    wz = torch.zeros(128, 2, 7, 7, device="cuda", requires_grad=True)
    F.conv2d(masks.bfloat16().float(), wz, None, stride=2, padding=3).backward(dy.float().view(r, 14, 14, 128).permute(0, 3, 1, 2))
    want = wz.grad.view(128, 98)
    scale = want.abs().max().item()
    assert (dw - want).abs().max().item() <= 2e-5 * scale * max(1.0, r ** 0.5)
    # the replaced route: dY^T x im2col on the tcgen05 GEMM
    col = ops.im2col_mask(masks, torch.bfloat16)
    ref = torch.empty(128, 104, device="cuda", dtype=torch.float32)
    ops.gemm(dy, col, ref, a_major=MAJOR_MN, b_major=MAJOR_MN)
    assert (dw - ref[:, :98]).abs().max().item() <= 2e-5 * scale * max(1.0, r ** 0.5)


@pytest.mark.parametrize("r,nv", [(1, 1), (37, 3), (300, 2)])
def test_pool_bn_bwd_matches_the_two_kernels(cuda_lib, r, nv):
    """Fused MaxPool + BatchNorm + ReLU backward vs nlv_maxpool_bwd + nlv_bn_bwd (dense fp32 gradient map)."""
    from nlvsgg_b200 import ops
    g = torch.Generator().manual_seed(r + 3)
    x = torch.relu(torch.randn(r * 196, 128, generator=g)).bfloat16().cuda()            # conv + ReLU output
    pair_video, seg196, bounds = _videos(r, nv)
    seg49 = (seg196 // 4).to(torch.int32)
    rm, rv = torch.zeros(128).cuda(), torch.ones(128).cuda()
    mean, var = ops.bn_stats(x, seg196, nv, 128, 0.01, rm, rv)
    w = torch.randn(128, generator=g).cuda()
    b = torch.randn(128, generator=g).cuda()
    y, arg, xmax = ops.bn_apply_maxpool(x, pair_video, mean, var, w, b, r, want_xmax=True)
    rows = torch.arange(r * 49, device="cuda").view(-1, 1)
    pairs, cell = rows // 49, rows % 49
    tap = arg.long()
    src = pairs * 196 + (2 * (cell // 7) - 1 + tap // 3) * 14 + (2 * (cell % 7) - 1 + tap % 3)
    assert torch.equal(xmax, x[src, torch.arange(128, device="cuda").view(1, -1)])        # the activation at every argmax
    dp = torch.randn(r * 49, 128, generator=g).cuda()
    dx, dw, db, cs = ops.pool_bn_bwd(dp, arg, x, xmax, pair_video, seg196, seg49, nv, mean, var, w, r)
    dense = ops.maxpool_bwd(dp, arg, r, 128, torch.float32)
    dx2, dw2, db2 = ops.bn_bwd(dense, x, None, seg196, pair_video, nv, mean, var, w, True, dx_dtype=torch.float32, gate_by_x=True, row_div=196)
    scale = dx2.abs().max().item()
    assert (dx.float() - dx2).abs().max().item() <= 2 ** -7 * scale                       # bf16 rounding of the stored gradient
    assert (dx.float() - dx2.bfloat16().float()).abs().max().item() <= 2 ** -6 * scale
    assert (dw - dw2).abs().max().item() <= 1e-4 * dw2.abs().max().item() + 1e-5
    assert (db - db2).abs().max().item() <= 1e-4 * db2.abs().max().item() + 1e-5
    want = dx2.double().sum(0)
    assert (cs.double() - want).abs().max().item() <= 1e-4 * want.abs().max().item() + 1e-4 * scale


@pytest.mark.parametrize("r", [1, 2, 3, 64, 777])
@pytest.mark.parametrize("out_dtype", [torch.float32, torch.bfloat16])
def test_conv3x3_dgrad_implicit_gemm(cuda_lib, r, out_dtype):
    """Data gradient of Conv2d(128, 256, 3, 1, 1) as one implicit tcgen05 GEMM (4-D TMA boxes, zero halo) vs torch and vs the
    column-gradient product + col2im it replaces."""
    from nlvsgg_b200 import ops
    from nlvsgg_b200._C import MAJOR_MN
    g = torch.Generator().manual_seed(r)
    w = (0.05 * torch.randn(256, 128, 3, 3, generator=g)).cuda()
    dy = torch.randn(r * 49, 256, generator=g).bfloat16().cuda()                 # NHWC rows
    wt = w.view(256, 128 * 9).t().contiguous().view(128, 9, 256).bfloat16().view(128, 2304).contiguous()   # [ci][tap][co]
    dx = ops.conv3x3_dgrad(dy, wt, r, 256, 128, out_dtype)
    want = torch.nn.grad.conv2d_input((r, 128, 7, 7), w.bfloat16().float(), dy.float().view(r, 7, 7, 256).permute(0, 3, 1, 2), padding=1)
    want = want.permute(0, 2, 3, 1).reshape(r * 49, 128)
    tol = 1e-2 if out_dtype == torch.bfloat16 else 2e-5
    assert (dx.float() - want).abs().max().item() <= tol * want.abs().max().item() + 1e-5
    # the replaced route
    w_op = w.permute(0, 2, 3, 1).reshape(256, 1152).bfloat16().contiguous()      # [co][tap][ci]
    dcol = torch.empty(r * 49, 1152, device="cuda", dtype=torch.bfloat16)
    ops.gemm(dy, w_op, dcol, b_major=MAJOR_MN)
    ref = ops.col2im_3x3(dcol, r, 7, 7, 128)
    assert (dx.float() - ref).abs().max().item() <= 1.5e-2 * ref.abs().max().item()       # that route rounds the column gradient to bf16
