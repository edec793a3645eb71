"""Host-side label / loss-weight arrays (nlvsgg_b200/model.py:label_arrays) without a GPU: a weighted sum of per-row
losses built from those arrays must equal the mean over videos of the reference-style per-video loss
(tools/train_STTran.py:143-189 as restated in oracle/model.py:training_loss) for random predictions."""
import numpy as np
import torch
import torch.nn.functional as F

from nlvsgg_b200 import model as M, synth


def _weighted_loss(arr, obj_logits, logits26, labels):
    la, wa = torch.from_numpy(arr["lab_att"]), torch.from_numpy(arr["lab_w_att"])
    ce_obj = F.cross_entropy(obj_logits, labels, reduction="none")
    loss = (ce_obj * torch.from_numpy(arr["lab_w_obj"])).sum()
    ce_att = F.cross_entropy(logits26[:, :3], la.clamp(min=0), reduction="none")
    loss = loss + (ce_att * wa).sum()
    for bits, w, sl, ncls in ((arr["lab_spa_bits"], arr["lab_w_spa"], slice(3, 9), 6), (arr["lab_con_bits"], arr["lab_w_con"], slice(9, 26), 17)):
        tgt = torch.from_numpy(((bits[:, None].astype(np.int64) >> np.arange(ncls)) & 1).astype(np.float32))
        bce = F.binary_cross_entropy_with_logits(logits26[:, sl], tgt, reduction="none").sum(1)
        loss = loss + (bce * torch.from_numpy(w)).sum()
    return loss


def test_label_arrays_reproduce_the_mean_of_per_video_reference_losses():
    from oracle import model as omodel
    entries = [synth.synth_video(900 + i, 5 + 2 * i, 5, "sgdet", with_gt=False, empty_frame_prob=0.2 * (i % 2))[0] for i in range(4)]
    entries[2]["attention_gt"][1] = []          # a pair without an attention label drops out of that mean (:150-151)
    entries[1]["spatial_gt"][0] = []            # same for the multi-label heads (:163-166)
    hb = M.collate(entries, "sgdet")
    arr = M.label_arrays(hb)
    g = torch.Generator().manual_seed(0)
    N, R = sum(hb.n_boxes), sum(hb.n_pairs)
    obj_logits, logits26 = torch.randn(N, 37, generator=g), torch.randn(R, 26, generator=g)
    got = _weighted_loss(arr, obj_logits, logits26, hb.labels)
    want, b0, r0 = 0.0, 0, 0
    for e, nb, nr in zip(entries, hb.n_boxes, hb.n_pairs):
        l26 = logits26[r0:r0 + nr]
        pred = {"distribution": obj_logits[b0:b0 + nb], "attention_distribution": l26[:, :3],
                "spatial_distribution": torch.sigmoid(l26[:, 3:9]), "contacting_distribution": torch.sigmoid(l26[:, 9:])}
        want = want + omodel.training_loss(pred, e, "sgdet") / len(entries)
        b0 += nb; r0 += nr
    assert abs(float(got) - float(want)) <= 2e-6 * abs(float(want))
    assert arr["lab_att"][hb.n_pairs[0] + hb.n_pairs[1] + 1] == -1 and arr["lab_w_att"][hb.n_pairs[0] + hb.n_pairs[1] + 1] == 0
