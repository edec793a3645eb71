"""Union-mask rasteriser and float64 IoU kernels vs the oracle (bit-exact)."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _pairs(n, seed):
    rng = np.random.default_rng(seed)
    b = rng.uniform(0, 400, (n, 8)).astype(np.float32)
    b[:, 2:4] = b[:, 0:2] + rng.uniform(1, 200, (n, 2)).astype(np.float32)
    b[:, 6:8] = b[:, 4:6] + rng.uniform(1, 200, (n, 2)).astype(np.float32)
    return b


@pytest.mark.parametrize("n,seed", [(1, 0), (97, 1), (4000, 2)])
def test_draw_union_boxes_bit_exact(cuda_lib, n, seed):
    from nlvsgg_b200 import ops
    from oracle import cref
    b = _pairs(n, seed)
    want = cref.draw_union_boxes(b, 27)
    got = ops.draw_union_boxes(torch.from_numpy(b).cuda(), 27).cpu().numpy()
    assert np.array_equal(got, want)
    want_off = want - np.float32(0.5)
    got_off = ops.draw_union_boxes(torch.from_numpy(b).cuda(), 27, -0.5).cpu().numpy()
    assert np.array_equal(got_off, want_off)


def test_draw_union_boxes_golden(cuda_lib):
    """Fixture produced by the reference's own Cython build (oracle/make_golden.py)."""
    import os
    from nlvsgg_b200 import ops
    path = os.path.join(os.path.dirname(__file__), "golden", "native_draw_union_boxes.npz")
    z = np.load(path)
    got = ops.draw_union_boxes(torch.from_numpy(z["box_pairs"]).cuda(), 27).cpu().numpy()
    assert np.array_equal(got, z["out"])


def test_draw_union_boxes_empty(cuda_lib):
    from nlvsgg_b200 import ops
    out = ops.draw_union_boxes(torch.zeros(0, 8, device="cuda"), 27)
    assert out.shape == (0, 2, 27, 27)


def test_union_mask_pairs_matches_gather(cuda_lib):
    from nlvsgg_b200 import ops, synth
    from oracle import cref
    entry, _ = synth.synth_video(3, 20, 6, "sgdet", draw_fn=cref.draw_union_boxes)
    got = ops.union_mask_pairs(entry["boxes"].cuda(), entry["pair_idx"].cuda(), 27, -0.5).cpu()
    assert torch.equal(got, entry["spatial_masks"])


@pytest.mark.parametrize("n,k", [(1, 1), (50, 70), (300, 17)])
def test_bbox_overlaps_bit_exact(cuda_lib, n, k):
    from nlvsgg_b200 import ops
    from oracle import cref
    rng = np.random.default_rng(n + k)
    x = rng.uniform(0, 100, (n, 4)); x[:, 2:] += x[:, :2]
    y = rng.uniform(0, 100, (k, 4)); y[:, 2:] += y[:, :2]
    x = x.astype(np.float32).astype(np.float64)  # evaluator rounds boxes to f32 first (evaluation_recall.py:765)
    want = cref.bbox_overlaps(x, y)
    got = ops.bbox_overlaps(torch.from_numpy(x).cuda(), torch.from_numpy(y).cuda()).cpu().numpy()
    assert np.array_equal(got, want)
