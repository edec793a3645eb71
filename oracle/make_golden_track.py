"""TEST INFRASTRUCTURE — tracking / RoIAlign / NMS golden vectors from the reference (build container only)."""
import copy
import os

import numpy as np
import torch

from nlvsgg_b200 import synth
from oracle import ref_harness as H

GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def track_entry(seed, frames=8, k=5, frame_stride=1, drift=6.0):
    """A video in which objects persist across frames (slowly drifting boxes, stable features) so tracks form."""
    g = torch.Generator().manual_seed(seed)
    n_obj = k
    base_box = torch.rand(n_obj, 4, generator=g)
    base_box = torch.stack((base_box[:, 0] * 250, base_box[:, 1] * 120, base_box[:, 0] * 250 + 40 + base_box[:, 2] * 120,
                            base_box[:, 1] * 120 + 30 + base_box[:, 3] * 90), 1)
    base_feat = torch.relu(torch.randn(n_obj, 2048, generator=g))
    labels0 = torch.randint(2, 37, (n_obj,), generator=g)
    labels0[0] = 1
    boxes, feats, dist, labels, gt = [], [], [], [], []
    for f in range(frames):
        present = torch.rand(n_obj, generator=g) > 0.25
        present[0] = True
        for o in range(n_obj):
            if not present[o]:
                continue
            b = base_box[o] + (torch.rand(4, generator=g) - 0.5) * drift + f * 1.5
            boxes.append([float(f)] + b.tolist())
            feats.append(base_feat[o] + 0.15 * torch.randn(2048, generator=g))
            d = torch.full((36,), 0.01); d[labels0[o] - 1] = 0.6 + 0.3 * float(torch.rand(1, generator=g))
            if float(torch.rand(1, generator=g)) < 0.15:      # occasional mis-classification
                d[int(torch.randint(0, 36, (1,), generator=g))] = 0.95
            dist.append(d); labels.append(int(labels0[o]))
        gt.append([{"person_bbox": np.zeros((1, 4), np.float32), "frame": "v.mp4/%06d.png" % (f * frame_stride)}])
    entry = {"boxes": torch.tensor(boxes, dtype=torch.float32), "features": torch.stack(feats), "distribution": torch.stack(dist),
             "labels": torch.tensor(labels, dtype=torch.int64)}
    return entry, gt


def main(ref=None):
    ref = ref or H.load_reference()
    matcher = ref.matcher.HungarianMatcher(0.5, 1, 1, 0.5)          # tools/train_DSG_DETR.py:113
    cases = []
    for name, seed, frames, k, stride in (("track_a", 81, 8, 5, 1), ("track_gap", 82, 10, 6, 30), ("track_b", 83, 14, 8, 3)):
        entry, gt = track_entry(seed, frames, k, stride)
        out = {}
        for task in ("sgcls", "sgdet", "predcls"):
            e = {kk: v.clone() for kk, v in entry.items()}
            ref.track.get_sequence(e, gt, matcher, (480, 270), task)
            out[task] = [t.long().tolist() if torch.is_tensor(t) and t.numel() else [] for t in e["indices"]]
        torch.save({"name": name, "seed": seed, "frames": frames, "k": k, "stride": stride, "indices": out},
                   os.path.join(GOLDEN, name + ".pt"))
        print("wrote", name, {t: len(v) for t, v in out.items()})
    # one matcher cost matrix
    entry, gt = track_entry(90, 3, 7, 1)
    f0 = entry["boxes"][:, 0] == 0; f1 = entry["boxes"][:, 0] == 1
    xywh = lambda b: torch.cat((b[:, :2], b[:, 2:] - b[:, :2]), 1) / torch.tensor([[480., 270., 480., 270.]])
    o = {"boxes": xywh(entry["boxes"][f1, 1:]), "features": entry["features"][f1], "dists": entry["distribution"][f1]}
    t = {"boxes": xywh(entry["boxes"][f0, 1:]), "features": entry["features"][f0], "dists": entry["distribution"][f0]}
    cd = ref.matcher.cost_matrix_torch(o["dists"], t["dists"]); cf = ref.matcher.cost_matrix_torch(o["features"], t["features"])
    ob, tb = ref.matcher.box_xywh_to_cxcywh(o["boxes"]), ref.matcher.box_xywh_to_cxcywh(t["boxes"])
    C = 0.5 * cd + 1 * cf + 1 * torch.cdist(ob, tb, p=1) + 0.5 * (-ref.matcher.generalized_box_iou(
        ref.matcher.box_cxcywh_to_xyxy(ob), ref.matcher.box_cxcywh_to_xyxy(tb)))
    r, c, c1, c2 = matcher(o, t)
    torch.save({"out": o, "tgt": t, "C": C, "cost_dist": cd, "cost_feat": cf, "row": np.asarray(r), "col": np.asarray(c)},
               os.path.join(GOLDEN, "track_cost.pt"))
    print("wrote track_cost", tuple(C.shape))


if __name__ == "__main__":
    main()
