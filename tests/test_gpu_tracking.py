"""Tracking cost kernel, get_sequence drop-in, NMS and RoIAlign kernels vs the oracle and reference golden vectors."""
import numpy as np
import pytest
import torch

from tests import golden_util as G

pytestmark = pytest.mark.gpu


def test_track_cost_kernel_matches_oracle_and_reference(cuda_lib):
    from nlvsgg_b200.lib.matcher import HungarianMatcher, track_cost
    from oracle import tracking as ot
    z = G.load_case("track_cost")
    o = {k: v.cuda() for k, v in z["out"].items()}
    t = {k: v.cuda() for k, v in z["tgt"].items()}
    C, cd, cf = track_cost(o["boxes"], t["boxes"], o["features"], t["features"], o["dists"], t["dists"], 0.5, 1.0, 1.0, 0.5)
    assert torch.allclose(C.cpu(), z["C"], rtol=1e-5, atol=2e-6)
    assert torch.allclose(cd.cpu(), z["cost_dist"], rtol=1e-5, atol=2e-6) and torch.allclose(cf.cpu(), z["cost_feat"], rtol=1e-5, atol=2e-6)
    r, c, c1, c2 = HungarianMatcher(0.5, 1, 1, 0.5)(o, t)
    assert r.tolist() == z["row"].tolist() and c.tolist() == z["col"].tolist()      # integer assignment: exact
    # larger random problem against the oracle restatement
    g = torch.Generator().manual_seed(3)
    mk = lambda n: {"boxes": torch.rand(n, 4, generator=g) * 0.4 + 0.05, "features": torch.relu(torch.randn(n, 2048, generator=g)),
                    "dists": torch.softmax(torch.randn(n, 36, generator=g) * 3, 1)}
    a, b = mk(20), mk(40)
    Cw, cdw, cfw = ot.matcher_cost(a, b)
    C, cd, cf = track_cost(a["boxes"].cuda(), b["boxes"].cuda(), a["features"].cuda(), b["features"].cuda(), a["dists"].cuda(),
                           b["dists"].cuda(), 0.5, 1.0, 1.0, 0.5)
    assert torch.allclose(C.cpu(), Cw, rtol=1e-5, atol=5e-6)


@pytest.mark.parametrize("name", ["track_a", "track_gap", "track_b"])
@pytest.mark.parametrize("task", ["sgcls", "sgdet", "predcls"])
def test_get_sequence_matches_reference(cuda_lib, name, task):
    from nlvsgg_b200.lib.matcher import HungarianMatcher
    from nlvsgg_b200.lib.track import get_sequence
    from oracle.make_golden_track import track_entry
    case = G.load_case(name)
    entry, gt = track_entry(case["seed"], case["frames"], case["k"], case["stride"])
    e = {k: v.cuda() for k, v in entry.items()}
    get_sequence(e, gt, HungarianMatcher(0.5, 1, 1, 0.5), (480, 270), task)
    got = [t.long().cpu().tolist() if t.numel() else [] for t in e["indices"]]
    assert got == case["indices"][task]                         # track assignment: bit-exact


def test_device_lsap_is_scipy(cuda_lib):
    """nlv_lsap vs scipy.optimize.linear_sum_assignment (the call of lib/matcher.py:147-149): identical assignments on
    rectangular problems either way round, including integer-valued costs full of ties (the tie rules are scipy's)."""
    from scipy.optimize import linear_sum_assignment as scipy_lsa
    from nlvsgg_b200.lib.matcher import linear_sum_assignment
    rng = np.random.default_rng(11)
    shapes = [(1, 1), (1, 7), (7, 1), (5, 5), (20, 40), (40, 20), (33, 33), (64, 97), (130, 70), (3, 300), (257, 256)]
    for n, m in shapes:
        for kind in ("float", "ties", "coarse"):
            if kind == "float":
                c = rng.standard_normal((n, m)).astype(np.float32)
            elif kind == "ties":
                c = rng.integers(0, 4, (n, m)).astype(np.float32)
            else:
                c = (rng.integers(0, 50, (n, m)) / 8.0).astype(np.float32)
            want_r, want_c = scipy_lsa(c)
            got_r, got_c = linear_sum_assignment(torch.from_numpy(c).cuda())
            assert got_r.cpu().tolist() == want_r.tolist() and got_c.cpu().tolist() == want_c.tolist(), (n, m, kind)
    r, c = linear_sum_assignment(torch.zeros(4, 0).cuda())
    assert r.numel() == 0 and c.numel() == 0


def test_track_videos_batch_equals_single(cuda_lib):
    """Several videos in one launch (one CTA each) give the clusters of the per-video calls."""
    from nlvsgg_b200.lib.track import frame_number, track_videos
    from oracle.make_golden_track import track_entry
    vids = []
    for name in ("track_a", "track_gap", "track_b"):
        case = G.load_case(name)
        entry, gt = track_entry(case["seed"], case["frames"], case["k"], case["stride"])
        vids.append((entry["boxes"].cuda(), entry["features"].cuda(), entry["distribution"].cuda(), [frame_number(a[0]["frame"]) for a in gt]))
    w = (0.5, 1.0, 1.0, 0.5)
    single = [track_videos([b], [f], [d], [k], (480, 270), w)[0] for b, f, d, k in vids]
    batch = track_videos([v[0] for v in vids], [v[1] for v in vids], [v[2] for v in vids], [v[3] for v in vids], (480, 270), w)
    for a, b in zip(single, batch):
        assert torch.equal(a, b)


@pytest.mark.parametrize("n,thr", [(1, 0.5), (64, 0.4), (200, 0.6), (513, 0.3)])
def test_nms_matches_oracle(cuda_lib, n, thr):
    from nlvsgg_b200.lib.roi_layers import nms
    from oracle import cref
    rng = np.random.default_rng(n)
    d = rng.uniform(0, 300, (n, 4)).astype(np.float32)
    d[:, 2:] = d[:, :2] + rng.uniform(5, 120, (n, 2)).astype(np.float32)
    s = rng.uniform(0, 1, n).astype(np.float32)
    for strict in (True, False):
        want = cref.nms(d, s, thr, strict=strict)
        got = nms(torch.from_numpy(d).cuda(), torch.from_numpy(s).cuda(), thr, strict=strict).cpu().numpy()
        assert np.array_equal(got, want)
    assert nms(torch.zeros(0, 4).cuda(), torch.zeros(0).cuda(), 0.5).numel() == 0


def test_roi_align_forward_bit_exact_and_backward(cuda_lib):
    import torchvision
    from nlvsgg_b200.lib.roi_layers import ROIAlign
    from oracle import cref
    rng = np.random.default_rng(0)
    inp = rng.standard_normal((3, 16, 38, 67)).astype(np.float32)
    rois = np.array([[0, 10, 20, 300, 200], [1, 0, 0, 1071, 607], [0, 5, 5, 6, 6], [2, -50, -50, 2000, 900], [1, 100.5, 33.3, 420.7, 199.9]],
                    np.float32)
    rois = np.concatenate([rois, np.column_stack((rng.integers(0, 3, 40), rng.uniform(0, 500, 40), rng.uniform(0, 300, 40),
                                                  rng.uniform(500, 1000, 40), rng.uniform(300, 600, 40))).astype(np.float32)])
    for sr in (0, 2):
        want = cref.roi_align_forward(inp, rois, 1 / 16., 7, 7, sr)
        x = torch.from_numpy(inp).cuda().requires_grad_(True)
        out = ROIAlign((7, 7), 1 / 16., sr)(x, torch.from_numpy(rois).cuda())
        assert np.array_equal(out.detach().cpu().numpy(), want)             # bit-identical to the reference CPU kernel
        gout = torch.from_numpy(rng.standard_normal(want.shape).astype(np.float32)).cuda()
        out.backward(gout)
        xc = torch.from_numpy(inp).requires_grad_(True)
        torchvision.ops.roi_align(xc, torch.from_numpy(rois), (7, 7), 1 / 16., sr, aligned=False).backward(gout.cpu())
        assert torch.allclose(x.grad.cpu(), xc.grad, rtol=1e-4, atol=1e-5)


def test_cython_surface_dropins(cuda_lib):
    from nlvsgg_b200.lib.draw_rectangles.draw_rectangles import draw_union_boxes
    from nlvsgg_b200.lib.fpn.box_intersections_cpu.bbox import bbox_overlaps
    z = np.load(G.GOLDEN + "/native_draw_union_boxes.npz")
    assert np.array_equal(draw_union_boxes(z["box_pairs"], 27), z["out"])
    z = np.load(G.GOLDEN + "/native_bbox_overlaps.npz")
    assert np.array_equal(bbox_overlaps(z["boxes"], z["query"]), z["out"])
