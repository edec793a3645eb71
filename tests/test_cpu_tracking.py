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
