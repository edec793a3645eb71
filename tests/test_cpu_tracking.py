"""Oracle tracking cost vs the reference matcher's numbers (golden)."""
import torch

from oracle import tracking as ot
from tests import golden_util as G


def test_oracle_cost_matches_reference_matcher():
    z = G.load_case("track_cost")
    C, cd, cf = ot.matcher_cost(z["out"], z["tgt"], 0.5, 1.0, 1.0, 0.5)
    assert torch.allclose(C, z["C"], rtol=1e-6, atol=1e-6)
    assert torch.allclose(cd, z["cost_dist"], rtol=1e-6, atol=1e-6) and torch.allclose(cf, z["cost_feat"], rtol=1e-6, atol=1e-6)
    from scipy.optimize import linear_sum_assignment
    r, c = linear_sum_assignment(C)
    assert r.tolist() == z["row"].tolist() and c.tolist() == z["col"].tolist()


def test_lsap_restatement_is_scipy():
    """oracle.tracking.lsap (the algorithm csrc/track.cu:lsap_warp follows) against scipy's linear_sum_assignment, including
    rectangular problems both ways round and integer costs full of ties."""
    import numpy as np
    from scipy.optimize import linear_sum_assignment
    from oracle.tracking import lsap
    rng = np.random.default_rng(5)
    for n, m in [(1, 1), (1, 6), (6, 1), (5, 5), (9, 14), (14, 9), (20, 20), (12, 31)]:
        for kind in range(3):
            c = (rng.standard_normal((n, m)) if kind == 0 else rng.integers(0, 3, (n, m)) if kind == 1 else rng.integers(0, 40, (n, m)) / 8.0)
            c = c.astype(np.float32)
            r0, c0 = linear_sum_assignment(c)
            r1, c1 = lsap(c)
            assert r0.tolist() == list(r1) and c0.tolist() == list(c1), (n, m, kind)
