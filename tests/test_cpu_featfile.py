"""Packed per-video feature files (nlvsgg_b200/featfile.py): write -> read round trip on the host."""
import numpy as np
import torch

from nlvsgg_b200 import featfile as FF, model as M, synth


def _entries():
    return [synth.synth_video(900 + i, 4 + 2 * i, 5, "sgdet", empty_frame_prob=0.2 if i == 1 else 0.0)[0] for i in range(3)]


def _unpack_numpy(hb):
    """numpy restatement of csrc/util.cu:union_unpack_kernel."""
    bm = hb.union_bitmap.numpy().view(np.uint64)
    bits = np.unpackbits(bm.view(np.uint8).reshape(bm.shape[0], 256), axis=1, bitorder="little").astype(bool)
    off = hb.union_off.numpy().view(np.uint32).astype(np.int64)
    dense = np.zeros(bits.shape, dtype=np.uint16)
    if hb.union_rows == 3:      # 12-bit values: low byte + 4-bit code on top of the row's base high byte (union_unpack12_kernel)
        lo, hx, base = hb.union_feat.numpy(), hb.union_hx.numpy(), hb.union_base.numpy()
        assert off[0] == 0 and np.all(np.diff(off) - bits.sum(1) >= 0) and int((np.diff(off) - bits.sum(1)).sum()) <= len(hb.n_pairs)
        rows, _ = np.nonzero(bits)
        k = np.arange(len(rows)) - np.repeat(np.cumsum(bits.sum(1)) - bits.sum(1), bits.sum(1))     # index of the value inside its row
        gi = off[rows] + k
        code = (hx[gi >> 1] >> (4 * (gi & 1))) & 15
        dense[bits] = ((base[rows].astype(np.uint16) + code) << 8) | lo[gi]
        if hb.union_exc_pos is not None:        # values outside their row's window (union_patch_kernel)
            dense.reshape(-1)[hb.union_exc_pos.numpy().view(np.uint32)] = hb.union_exc_val.numpy().view(np.uint16)
        return dense
    vals = hb.union_feat.view(torch.int16).numpy().view(np.uint16)
    dense[bits] = vals[:int(bits.sum())]
    assert off[0] == 0 and np.array_equal(np.diff(off), bits.sum(1))
    return dense


import pytest


@pytest.mark.parametrize("pack12", [True, False])
def test_round_trip_matches_collate(tmp_path, pack12):
    entries = _entries()
    paths = FF.write_videos(str(tmp_path), entries, pack12=pack12)
    hb = FF.Loader(pin=False).load(paths)
    ref = M.collate(entries, "sgdet")
    assert hb.n_boxes == ref.n_boxes and hb.n_pairs == ref.n_pairs
    assert all(np.array_equal(a, b) for a, b in zip(hb.frame_ids, ref.frame_ids))
    assert torch.equal(hb.boxes, ref.boxes) and torch.equal(hb.labels, ref.labels) and torch.equal(hb.scores, ref.scores)
    assert torch.equal(hb.pair_idx, ref.pair_idx)
    assert torch.equal(hb.features, ref.features.bfloat16())
    # the distribution of the synthetic producer is a create_dis one: stored as (confidence, class) and rebuilt bit for bit
    assert hb.distribution is None
    rebuilt = hb.dist_other[:, None].expand(-1, 36).clone()
    rebuilt[torch.arange(len(hb.dist_idx)), hb.dist_idx.long()] = hb.dist_conf
    assert torch.equal(rebuilt, ref.distribution)
    # union features: channels-last, zero-suppressed, lossless in bf16
    assert hb.union_rows == (3 if pack12 else 2)
    want = ref.union_feat.bfloat16().permute(0, 2, 3, 1).reshape(-1, 2048).contiguous().view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(_unpack_numpy(hb), want)
    # labels: the CSR form gives the same label tensors and loss weights as the python lists
    a, b = M.label_arrays(hb), M.label_arrays(ref)
    assert a.keys() == b.keys() and all(np.array_equal(a[k], b[k]) for k in a)
    assert M.input_bytes(hb) < (0.27 if pack12 else 0.35) * M.input_bytes(ref)          # vs the fp32 NCHW entry contract


def test_12_bit_values_exceptions_and_fallback(tmp_path):
    """Values outside a row's 16-step window (vanishing magnitudes, negative values) travel as exceptions; a video with too many of
    them keeps 16-bit values, and write_videos then writes the whole set that way."""
    def want_of(e):
        return e["union_feat"].bfloat16().permute(0, 2, 3, 1).reshape(-1, 2048).contiguous().view(torch.int16).numpy().view(np.uint16)
    e = _entries()[0]
    e["union_feat"][0, 5, 3, 3] = -1.5                       # a negative value
    e["union_feat"][1, 7, 0, 0] = 1e-30                      # 100 binades below the row's maximum
    e["union_feat"][1, 8, 0, 0] = 2.0 ** -20                 # inside the window
    e["union_feat"][2, 9, 1, 1] = -0.0
    p = str(tmp_path / "v.nlvf")
    meta = FF.write_video(p, e)
    assert meta["union"] == "sparse12" and meta["union_nexc"] == 3
    hb = FF.Loader(pin=False).load([p])
    assert hb.union_rows == 3 and hb.union_exc_pos.numel() == 3 and np.array_equal(_unpack_numpy(hb), want_of(e))
    # exceptions of the second video of a batch are re-based to batch-global rows
    e2 = _entries()[1]
    e2["union_feat"][3, 100, 2, 2] = -2.0
    paths = FF.write_videos(str(tmp_path / "two"), [e, e2])
    hb = FF.Loader(pin=False).load(paths)
    assert hb.union_exc_pos.numel() == 4 and np.array_equal(_unpack_numpy(hb), np.concatenate([want_of(e), want_of(e2)]))
    # mixed signs everywhere: half of the values are exceptions -> 16-bit values, for the whole set
    g = torch.Generator().manual_seed(0)
    e3 = _entries()[2]
    e3["union_feat"] = torch.randn(e3["union_feat"].shape, generator=g) * (torch.rand(e3["union_feat"].shape, generator=g) > 0.5)
    assert FF.write_video(p, e3)["union"] == "sparse"
    paths = FF.write_videos(str(tmp_path / "three"), [e, e3])
    assert [FF.read_header(q)["union"] for q in paths] == ["sparse", "sparse"]
    hb = FF.Loader(pin=False).load(paths)
    assert hb.union_rows == 2 and np.array_equal(_unpack_numpy(hb), np.concatenate([want_of(e), want_of(e3)]))


def test_dense_union_and_full_distribution_fallbacks(tmp_path):
    e = _entries()[0]
    e["union_feat"] = e["union_feat"] + 1.0            # no zeros -> dense rows are smaller than bitmap + values
    e["distribution"] = torch.softmax(torch.randn(e["boxes"].shape[0], 36, generator=torch.Generator().manual_seed(1)), 1)
    p = str(tmp_path / "v.nlvf")
    meta = FF.write_video(p, e)
    assert meta["union"] == "dense" and meta["dist"] == "full"
    hb = FF.Loader(pin=False).load([p])
    assert hb.union_rows == 1 and hb.union_bitmap is None
    assert torch.equal(hb.union_feat.view(-1, 49, 2048), e["union_feat"].bfloat16().permute(0, 2, 3, 1).reshape(-1, 49, 2048))
    assert torch.equal(hb.distribution, e["distribution"])


def test_negative_zero_is_kept(tmp_path):
    e = _entries()[0]
    e["union_feat"][0, 0, 0, 0] = -0.0
    e["union_feat"][0, 1, 0, 0] = 0.0
    p = str(tmp_path / "v.nlvf")
    FF.write_video(p, e, sparse=True)
    hb = FF.Loader(pin=False).load([p])
    d = _unpack_numpy(hb)
    assert d[0, 0] == 0x8000 and d[0, 1] == 0
