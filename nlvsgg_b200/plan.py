"""Host-side batch descriptors (vectorised numpy) + one pinned, asynchronous upload.

Replaces the four per-frame python loops of lib/transformer_wk.py:140-171,199-215 (pad / mask / window gather /
'latter' scatter, each with several device syncs per frame) by index arrays computed once per batch on the host
from the frame ids, shipped to the device in a single cudaMemcpyAsync from pinned memory (a pageable copy would
serialise the host with the stream and stall the pipeline every step)."""
from __future__ import annotations

from typing import Dict, List, Optional

import numpy as np
import torch


_TORCH_DT = {np.dtype(np.int32): torch.int32, np.dtype(np.int64): torch.int64, np.dtype(np.float32): torch.float32,
             np.dtype(np.uint32): torch.int32}


class Stager:
    """Ring of reusable (pinned host buffer, device buffer) pairs for the per-step descriptor upload.

    Allocating pinned memory inside the step is what must not happen: cudaHostAlloc synchronises the device, and the
    caching host allocator cannot recycle a block whose copy is still queued behind the previous step's kernels.  A slot
    is reused `depth` uploads later, after its consumer (the step that read the device copy) has finished — that wait is
    the only back-pressure on a host that runs ahead of the GPU."""

    def __init__(self, device, depth: int = 4):
        self.device, self.depth, self.i = torch.device(device), depth, 0
        self.slots = [dict(host=None, dev=None, done=None) for _ in range(depth)]

    def next(self, nbytes: int):
        s = self.slots[self.i % self.depth]
        self.i += 1
        if s["done"] is not None:
            s["done"].synchronize()
            s["done"] = None
        if s["host"] is None or s["host"].numel() < nbytes:
            cap = int(nbytes * 1.5) + 4096
            s["host"] = torch.empty(cap, dtype=torch.uint8, pin_memory=self.device.type == "cuda")
            s["dev"] = torch.empty(cap, dtype=torch.uint8, device=self.device)
        return s


def pack_upload(arrays: Dict[str, np.ndarray], device, consumer_stream=None, stager: Optional[Stager] = None) -> Dict[str, torch.Tensor]:
    """All arrays -> one pinned byte buffer -> one async H2D copy -> typed device views.  With a `stager` no memory is
    allocated; the returned dict carries the slot under "_slot" (record slot["done"] after the consumer is enqueued)."""
    offs, total = {}, 0
    for k, a in arrays.items():
        total = (total + 15) // 16 * 16
        offs[k] = total
        total += a.nbytes
    total = max(total, 16)
    slot = None
    if stager is not None:
        slot = stager.next(total)
        host, dev = slot["host"], slot["dev"]
    else:
        host = torch.empty(total, dtype=torch.uint8, pin_memory=torch.cuda.is_available())
    hv = host.numpy()
    for k, a in arrays.items():
        if a.nbytes:
            hv[offs[k]:offs[k] + a.nbytes] = np.ascontiguousarray(a).view(np.uint8).reshape(-1)
    if slot is not None:
        dev[:total].copy_(host[:total], non_blocking=True)
    else:
        dev = host.to(device, non_blocking=True)
        if consumer_stream is not None and dev.is_cuda:
            dev.record_stream(consumer_stream)   # issued on a copy stream, read on the compute stream
    out = {}
    for k, a in arrays.items():
        t = dev[offs[k]:offs[k] + a.nbytes]
        tdt = _TORCH_DT[a.dtype]
        out[k] = t.view(tdt).reshape(a.shape) if a.nbytes else torch.empty(a.shape, dtype=tdt, device=device)
    out["_pinned_keepalive"] = host
    out["_slot"] = slot
    return out


def work_items(seg_start: np.ndarray, seg_len: np.ndarray, seg_pad: Optional[np.ndarray] = None) -> np.ndarray:
    """int4 work list for the attention kernels: 16 rows of one segment per item.  seg_pad: padded keys of each segment
    (4th field; only read by the additive-int-mask compatibility mode, lib/transformer_wk.py:154)."""
    seg_start = np.asarray(seg_start, dtype=np.int64)
    seg_len = np.asarray(seg_len, dtype=np.int64)
    nblk = (seg_len + 15) // 16
    total = int(nblk.sum())
    if total == 0:
        return np.zeros((0, 4), dtype=np.int32)
    seg = np.repeat(np.arange(len(seg_len)), nblk)
    first = np.repeat(np.cumsum(nblk) - nblk, nblk)
    q0 = (np.arange(total) - first) * 16
    out = np.zeros((total, 4), dtype=np.int32)
    out[:, 0], out[:, 1], out[:, 2] = seg_start[seg], seg_len[seg], q0
    if seg_pad is not None:
        out[:, 3] = np.asarray(seg_pad, dtype=np.int64)[seg]
    return out


def long_first(work: np.ndarray):
    """Work list with the items of segments longer than 16 rows in front (stable), and their count: the two-kernel attention
    backward is launched over those alone (nlv_attn_bwd_sorted)."""
    is_long = work[:, 1] > 16
    n_long = int(is_long.sum())
    if n_long == 0 or n_long == len(work):
        return work, n_long
    return np.ascontiguousarray(work[np.argsort(~is_long, kind="stable")]), n_long


def _ranges(starts: np.ndarray, lens: np.ndarray) -> np.ndarray:
    """Concatenation of arange(starts[i], starts[i]+lens[i])."""
    total = int(lens.sum())
    if total == 0:
        return np.zeros(0, dtype=np.int64)
    first = np.repeat(np.cumsum(lens) - lens, lens)
    return np.repeat(starts, lens) + (np.arange(total) - first)


class Plan:
    """Everything the kernels need to know about the batch structure."""

    def __init__(self, n_boxes: List[int], frame_ids: List[np.ndarray], device, obj_class: Optional[np.ndarray] = None,
                 subj_box: Optional[np.ndarray] = None, dsg: bool = False, dsg_pos_by_rank: bool = True,
                 extra: Optional[Dict[str, np.ndarray]] = None, consumer_stream=None, stager: Optional[Stager] = None):
        nv = len(n_boxes)
        self.nv = nv
        n_pairs = np.asarray([len(f) for f in frame_ids], dtype=np.int64)
        n_boxes = np.asarray(n_boxes, dtype=np.int64)
        self.N, self.R = int(n_boxes.sum()), int(n_pairs.sum())
        box_seg = np.concatenate(([0], np.cumsum(n_boxes))).astype(np.int32)
        pair_seg = np.concatenate(([0], np.cumsum(n_pairs))).astype(np.int32)

        # global frame key: video-major, frame-minor
        fids = [np.asarray(f).astype(np.int64) for f in frame_ids]
        nfr = np.asarray([int(f[-1]) + 1 if len(f) else 0 for f in fids], dtype=np.int64)   # b = last frame id + 1
        for f in fids:
            assert len(f) == 0 or np.all(np.diff(f) >= 0), "im_idx must be sorted"
        fbase = np.concatenate(([0], np.cumsum(nfr)))
        gkey = np.concatenate([f + fbase[v] for v, f in enumerate(fids)]) if self.R else np.zeros(0, dtype=np.int64)
        F = int(fbase[-1])
        cnt = np.bincount(gkey, minlength=F) if self.R else np.zeros(F, dtype=np.int64)
        start = np.concatenate(([0], np.cumsum(cnt)))                    # first token row of every (video, frame)
        # ---- frames = segments of the spatial encoder
        nz = cnt > 0
        # padded keys of a frame = longest frame of its video (l, lib/transformer_wk.py:133) - its own pairs
        has = nfr > 0
        lmax = np.repeat(np.maximum.reduceat(cnt, fbase[:-1][has]), nfr[has]) if F else np.zeros(0, dtype=np.int64)
        lw, self.n_local_long = long_first(work_items(start[:-1][nz], cnt[nz], (lmax - cnt)[nz]))
        # ---- windows {j, j+1} inside each video (lib/transformer_wk.py:163-185): keep those with any token
        is_last = np.zeros(F, dtype=bool)
        is_last[fbase[1:][nfr > 0] - 1] = True
        j = np.nonzero(~is_last & ((cnt + np.concatenate((cnt[1:], [0]))) > 0))[0] if F else np.zeros(0, dtype=np.int64)
        n0, n1 = cnt[j], cnt[j + 1] if len(j) else cnt[j]
        wlen = n0 + n1
        wbase = np.cumsum(wlen) - wlen
        stream_src = _ranges(start[j], wlen)                              # token row copied into each stream row
        self.Mg = int(wlen.sum())
        pos_in_w = np.arange(self.Mg) - np.repeat(wbase, wlen)
        slot = (pos_in_w >= np.repeat(n0, wlen)).astype(np.int32)
        first_of_video = np.zeros(F, dtype=bool)
        first_of_video[fbase[:-1][nfr > 0]] = True
        takes = (slot == 1) | np.repeat(first_of_video[j], wlen)          # 'latter' + window 0's first half (:209-215)
        out_src = np.full(self.R, -1, dtype=np.int32)
        out_src[stream_src[takes]] = np.nonzero(takes)[0]
        out_inv = np.full(self.Mg, -1, dtype=np.int32)
        out_inv[takes] = stream_src[takes]
        order = np.argsort(stream_src, kind="stable")
        ss = stream_src[order]
        is_first = np.ones(self.Mg, dtype=bool)
        is_first[1:] = ss[1:] != ss[:-1]
        inv = np.full((self.R, 2), -1, dtype=np.int32)
        inv[ss[is_first], 0] = order[is_first]
        inv[ss[~is_first], 1] = order[~is_first]
        # videos without any window (single frame): output = spatial-encoder output (:187-188)
        passthrough = np.where(out_src < 0, np.arange(self.R), -1).astype(np.int32)
        self.has_passthrough = bool((passthrough >= 0).any())
        gw, self.n_glob_long = long_first(work_items(wbase, wlen))

        both_w = (inv >= 0).sum(1).astype(np.float32)          # windows a token appears in (0, 1 or 2) -> 1 / count
        both_w = np.where(both_w > 0, 1.0 / np.maximum(both_w, 1.0), 0.0).astype(np.float32)
        arrays = {
            "both_w": both_w, "box_seg": box_seg, "pair_seg": pair_seg, "seg196": (pair_seg.astype(np.int64) * 196).astype(np.int32),
            "seg49": (pair_seg.astype(np.int64) * 49).astype(np.int32),
            "local_work": lw, "glob_work": gw, "stream_src": stream_src.astype(np.int32), "stream_slot": slot,
            "inv": inv, "out_src": out_src, "out_inv": out_inv, "passthrough": passthrough,
        }
        if nv > 1:
            pair_row = np.repeat(np.arange(nv, dtype=np.int32), n_pairs)
            arrays.update(box_row=np.repeat(np.arange(nv, dtype=np.int32), n_boxes), pair_row=pair_row)
        self.n_local_work, self.n_glob_work = len(lw), len(gw)

        # ---- DSG-DETR: per-video, per-object-class sequences (lib/dsg_detr.py:545-559)
        self.dsg = dsg
        if dsg:
            vid = np.repeat(np.arange(nv), n_pairs)
            key = vid * 64 + obj_class.astype(np.int64)
            perm = np.argsort(key, kind="stable")
            ks = key[perm]
            new = np.ones(self.R, dtype=bool)
            new[1:] = ks[1:] != ks[:-1]
            s_start = np.nonzero(new)[0]
            s_len = np.diff(np.concatenate((s_start, [self.R])))
            if dsg_pos_by_rank:
                # rank of the row's subject box among the sequence's distinct subject boxes, laid out as
                # [0]*count0 + [1]*count1 ... (:553-556)
                sb = subj_box[perm].astype(np.int64)
                seq = np.repeat(np.arange(len(s_start)), s_len)
                o2 = np.lexsort((sb, seq))
                sb2, seq2 = sb[o2], seq[o2]
                chg = np.ones(self.R, dtype=bool)
                chg[1:] = (sb2[1:] != sb2[:-1]) | (seq2[1:] != seq2[:-1])
                rank_sorted = np.cumsum(chg) - 1
                rank_sorted = rank_sorted - np.repeat(rank_sorted[np.concatenate(([0], np.nonzero(seq2[1:] != seq2[:-1])[0] + 1))], s_len)
                pos = rank_sorted                      # already "sorted within the sequence" = the reference layout
            else:
                pos = np.arange(self.R) - np.repeat(s_start, s_len)
            iperm = np.empty(self.R, dtype=np.int64)
            iperm[perm] = np.arange(self.R)
            cw, self.n_cls_long = long_first(work_items(s_start, s_len))
            arrays.update(cls_perm=perm.astype(np.int32), cls_iperm=iperm.astype(np.int32), cls_pos=pos.astype(np.int32), cls_work=cw)
            self.n_cls_work = len(cw)
        if extra:
            arrays.update(extra)
        dev = pack_upload(arrays, device, consumer_stream, stager)
        self._keep = dev.pop("_pinned_keepalive")
        self._slot = dev.pop("_slot")
        for k, v in dev.items():
            setattr(self, k, v)
        if nv == 1:
            self.box_row = self.pair_row = None

    def consumed(self):
        """Call after the last kernel that reads the descriptors has been enqueued (staged uploads only)."""
        if self._slot is not None and torch.cuda.is_available():
            ev = torch.cuda.Event()
            ev.record()
            self._slot["done"] = ev
