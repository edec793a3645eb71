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


def test_12_bit_values_fall_back_when_a_row_spans_too_many_binades(tmp_path):
    """A row whose stored values span more than 16 high bytes (here: a negative value next to positive ones) keeps 16-bit values."""
    e = _entries()[0]
    e["union_feat"][0, 5, 3, 3] = -1.5
    p = str(tmp_path / "v.nlvf")
    meta = FF.write_video(p, e)
    assert meta["union"] == "sparse"
    hb = FF.Loader(pin=False).load([p])
    want = e["union_feat"].bfloat16().permute(0, 2, 3, 1).reshape(-1, 2048).contiguous().view(torch.int16).numpy().view(np.uint16)
    assert hb.union_rows == 2 and np.array_equal(_unpack_numpy(hb), want)
    # and a wide positive range inside one row
    e = _entries()[0]
    e["union_feat"][0, :, 0, 0] = torch.relu(e["union_feat"][0, :, 0, 0]) + 1.0
    e["union_feat"][0, 7, 0, 0] = 1e-30
    assert FF.write_video(p, e)["union"] == "sparse"
    e["union_feat"][0, 7, 0, 0] = 2.0 ** -20                      # 21 binades below: still inside 16 high bytes
    assert FF.write_video(p, e)["union"] == "sparse12"
    hb = FF.Loader(pin=False).load([p])
    want = e["union_feat"].bfloat16().permute(0, 2, 3, 1).reshape(-1, 2048).contiguous().view(torch.int16).numpy().view(np.uint16)
    assert np.array_equal(_unpack_numpy(hb), want)


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
