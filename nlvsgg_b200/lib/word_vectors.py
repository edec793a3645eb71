"""lib/word_vectors.py:15-35 stand-in: GloVe vectors when a local copy exists, else N(0,1) rows (what the reference
does for out-of-vocabulary tokens, word_vectors.py:18-19).  There is no network here."""
import os

import torch


def obj_edge_vectors(names, wv_type="glove.6B", wv_dir="data", wv_dim=200):
    path = os.path.join(os.environ.get("NLV_GLOVE_DIR", wv_dir), f"{wv_type}.{wv_dim}d.pt")
    g = torch.Generator().manual_seed(1234)
    vectors = torch.randn(len(names), wv_dim, generator=g)
    if os.path.exists(path):
        wv_dict, wv_arr, _ = torch.load(path)
        for i, token in enumerate(names):
            idx = wv_dict.get(token.split("/")[0])
            if idx is not None:
                vectors[i] = wv_arr[idx]
    return vectors
