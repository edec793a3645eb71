"""Host-side batch descriptors (nlvsgg_b200/plan.py) without a GPU: the window stream / 'latter' pick / work-item arrays are
checked against a literal per-frame restatement of lib/transformer_wk.py:140-215 and against structural invariants, for
videos with empty frames, skipped frame ids, a single frame, and for multi-video batches."""
import numpy as np
import pytest

from nlvsgg_b200.plan import Plan, work_items


def loop_restatement(f):
    """Per-frame python loops as the reference writes them: returns (stream_src, stream_slot, out_src) for one video."""
    f = np.asarray(f, dtype=np.int64)
    b = int(f[-1]) + 1 if len(f) else 0
    rows = [np.nonzero(f == t)[0] for t in range(b)]
    stream, slot, out_src = [], [], np.full(len(f), -1, dtype=np.int64)
    for j in range(b - 1):                                  # windows {j, j+1} (:163-171); empty ones are dropped (:175-185)
        if len(rows[j]) + len(rows[j + 1]) == 0:
            continue
        base = len(stream)
        stream += list(rows[j]) + list(rows[j + 1])
        slot += [0] * len(rows[j]) + [1] * len(rows[j + 1])
        if j == 0:                                          # frame 0 is read from the first half of window 0 (:209-211)
            out_src[rows[0]] = base + np.arange(len(rows[0]))
        out_src[rows[j + 1]] = base + len(rows[j]) + np.arange(len(rows[j + 1]))   # 'latter' (:212-215)
    return np.asarray(stream, dtype=np.int64), np.asarray(slot, dtype=np.int64), out_src


VIDEOS = {
    "dense": [0, 0, 0, 1, 1, 2, 2, 2, 2, 3],
    "first_frame_empty": [1, 1, 2, 3, 3],
    "gap_in_the_middle": [0, 0, 1, 4, 4, 5],
    "single_frame": [0, 0, 0],
    "single_late_frame": [3, 3],
    "long_tail": list(np.repeat(np.arange(12), [3, 1, 0, 2, 5, 0, 0, 1, 4, 2, 0, 3])),
}


@pytest.mark.parametrize("name", list(VIDEOS))
def test_window_stream_matches_the_per_frame_loops(name):
    f = np.asarray(VIDEOS[name], dtype=np.int64)
    p = Plan([len(f) + 1], [f], "cpu")
    stream, slot, out_src = loop_restatement(f)
    assert p.Mg == len(stream)
    assert np.array_equal(p.stream_src.numpy(), stream) and np.array_equal(p.stream_slot.numpy(), slot)
    assert np.array_equal(p.out_src.numpy(), out_src)
    # tokens of a video without any window pass the spatial-encoder output through (:187-188)
    assert np.array_equal(p.passthrough.numpy() >= 0, out_src < 0)
    assert p.has_passthrough == bool((out_src < 0).any())
    # inverse maps used by the backward pass
    inv = p.inv.numpy()
    for r in range(len(f)):
        assert sorted(x for x in inv[r] if x >= 0) == sorted(np.nonzero(stream == r)[0].tolist())
    oi = p.out_inv.numpy()
    assert np.array_equal(np.nonzero(oi >= 0)[0], np.sort(out_src[out_src >= 0]))
    assert all(stream[s] == oi[s] for s in np.nonzero(oi >= 0)[0])


def test_multi_video_batch_is_the_offset_concatenation_of_single_video_plans():
    vids = [np.asarray(VIDEOS[k], dtype=np.int64) for k in ("dense", "single_frame", "gap_in_the_middle", "first_frame_empty")]
    nb = [len(v) + 2 for v in vids]
    p = Plan(nb, vids, "cpu")
    roff = np.concatenate(([0], np.cumsum([len(v) for v in vids])))
    ss, sl, os_, soff = [], [], [], 0
    for i, v in enumerate(vids):
        a, b, c = loop_restatement(v)
        ss.append(a + roff[i]); sl.append(b); os_.append(np.where(c >= 0, c + soff, -1)); soff += len(a)
    assert np.array_equal(p.stream_src.numpy(), np.concatenate(ss)) and np.array_equal(p.stream_slot.numpy(), np.concatenate(sl))
    assert np.array_equal(p.out_src.numpy(), np.concatenate(os_))
    assert np.array_equal(p.pair_row.numpy(), np.repeat(np.arange(len(vids)), [len(v) for v in vids]))
    assert np.array_equal(p.box_seg.numpy(), np.concatenate(([0], np.cumsum(nb))))


def test_work_items_tile_every_segment_exactly_once():
    starts, lens = np.array([0, 5, 5, 40, 140]), np.array([5, 0, 35, 100, 17])
    w = work_items(starts, lens)
    covered = np.zeros(157, dtype=np.int64)
    for s, l, q0, _ in w:
        assert l > 0 and q0 % 16 == 0 and q0 < l
        covered[s + q0: s + min(q0 + 16, l)] += 1
    assert np.array_equal(covered, np.ones(157, dtype=np.int64))
    assert len(w) == sum((l + 15) // 16 for l in lens)


def test_dsg_class_sequences_and_subject_ranks():
    """lib/dsg_detr.py:545-559: per-video per-object-class sequences; position = rank of the subject box in the sequence."""
    f = np.array([0, 0, 0, 1, 1, 2, 2, 2], dtype=np.int64)
    obj_class = np.array([5, 9, 5, 5, 9, 9, 5, 5])
    subj_box = np.array([0, 0, 0, 4, 4, 7, 7, 7])
    p = Plan([10], [f], "cpu", obj_class=obj_class, subj_box=subj_box, dsg=True, dsg_pos_by_rank=True)
    perm, pos = p.cls_perm.numpy(), p.cls_pos.numpy()
    assert np.array_equal(obj_class[perm], np.sort(obj_class, kind="stable"))
    assert np.array_equal(perm, np.argsort(obj_class, kind="stable"))
    # class 5 rows: subjects 0,0,4,7,7 -> ranks 0,0,1,2,2 ; class 9 rows: subjects 0,4,7 -> 0,1,2
    assert pos.tolist() == [0, 0, 1, 2, 2, 0, 1, 2]
    assert np.array_equal(p.cls_iperm.numpy()[perm], np.arange(8))
    assert [tuple(x[:3]) for x in p.cls_work.numpy()] == [(0, 5, 0), (5, 3, 0)]


def test_long_first_orders_items_of_long_segments_in_front():
    """Work lists are ordered long-first (segments of more than 16 rows) so that the two-kernel attention backward is launched
    over those items only (nlv_attn_bwd_sorted); the order inside each class is kept."""
    import numpy as np
    from nlvsgg_b200.plan import long_first, work_items
    w = work_items(np.array([0, 5, 45, 50]), np.array([5, 40, 5, 17]))
    out, n_long = long_first(w)
    assert n_long == 3 + 2 and len(out) == len(w)
    assert (out[:n_long, 1] > 16).all() and (out[n_long:, 1] <= 16).all()
    assert out[:n_long].tolist() == [r for r in w.tolist() if r[1] > 16] and out[n_long:].tolist() == [r for r in w.tolist() if r[1] <= 16]
    short_only, n0 = long_first(work_items(np.array([0, 7]), np.array([7, 9])))
    assert n0 == 0 and short_only.tolist() == work_items(np.array([0, 7]), np.array([7, 9])).tolist()
