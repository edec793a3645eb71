#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json): STTran SGDet training step — forward, losses, backward, gradient
allreduce, clip + AdamW — over a batch of synthetic Action-Genome-shaped videos, frames/s.

    python bench.py --gpus N --steps K --warmup W                  # this repo (sm_100a kernels), config C2
    python bench.py --config {c1,c3,c4,c5} ...                      # the other BASELINE.json shapes
    python bench.py --impl reference --steps K --warmup W           # the reference formulation on the host CPU cores

One JSON line on stdout (rank 0).  See DESIGN.md §Measurement for every field.
"""
import argparse
import json
import os
import shutil
import subprocess
import sys
import tempfile
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

METRIC = "sttran_sgdet_train_frames_per_sec"
UNIT = "frames/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--config", default="c2", choices=["c1", "c2", "c3", "c4", "c5"],
                    help="BASELINE.json configs: c2 = the headline (64 videos/GPU, STTran sgdet train); c1 = predcls inference of one "
                         "video through the drop-in module; c3 = DSG-DETR, 8 videos/GPU; c4 = long videos (200 frames x 20 boxes); "
                         "c5 = Recall@K over the test-split shape")
    ap.add_argument("--videos", type=int, default=None, help="videos per GPU per step")
    ap.add_argument("--frames", type=int, default=None, help="mean frames per video (U{f-10..f+10})")
    ap.add_argument("--boxes", type=int, default=None, help="mean boxes per frame")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3", "fp32"])
    ap.add_argument("--arch", default=None, choices=["sttran", "dsg"])
    ap.add_argument("--input", default="packed", choices=["packed", "entry"],
                    help="packed = per-video feature files (featfile.py: bf16, channels-last, zero-suppressed union rows); "
                         "entry = the reference's fp32 NCHW entry tensors")
    ap.add_argument("--cpu-sample-videos", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-extras", action="store_true", help="skip the drop-in, parity-mode and fp32-entry legs (N=1 only anyway)")
    a = ap.parse_args()
    preset = {"c1": ("sttran", 1, 20, 6), "c2": ("sttran", 64, 30, 7), "c3": ("dsg", 8, 30, 7), "c4": ("sttran", 8, 200, 20),
              "c5": ("sttran", 1737, 31, 6)}[a.config]
    a.arch = a.arch or preset[0]
    a.videos = a.videos or preset[1]
    a.frames = a.frames or preset[2]
    a.boxes = a.boxes or preset[3]
    return a


def workload_name(a):
    return (f"{'STTran' if a.arch == 'sttran' else 'DSG-DETR'} SGDet training step, {a.videos} synthetic AG videos/GPU "
            f"(~{a.frames} frames, ~{a.boxes} VinVL 2048-d boxes/frame), {a.precision}")


def make_videos(a, rank, n, draw_fn=None, with_gt=False):
    """The STRUCTURE of video i (frames, boxes per frame, pairs, labels) is the same on every rank; its feature tensors are seeded
    per rank.  Every rank therefore steps through exactly the same number of frames and pairs (equal-size shards, as a
    length-bucketed sampler would hand out): the max-over-ranks step time measures communication, not load imbalance."""
    from nlvsgg_b200 import synth
    g = torch.Generator().manual_seed(777)
    out = []
    for i in range(n):
        if a.config == "c4":
            frames = a.frames
        else:
            frames = int(torch.randint(max(2, a.frames - 10), a.frames + 11, (1,), generator=g))
        e, _ = synth.synth_video(i, frames, a.boxes, "sgdet", draw_fn=draw_fn, with_gt=with_gt,
                                 fixed_boxes=a.boxes if a.config == "c4" else None, content_seed=100000 * rank + i + 1)
        out.append(e)
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False
        # NVML in-process (initialised here, before the timed region): a query is microseconds and takes no driver-wide lock,
        # whereas forking nvidia-smi every 100 ms showed up as multi-millisecond stalls of the launching thread
        self.nvml = self.handle = None
        try:
            import pynvml
            pynvml.nvmlInit()
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            phys = int(vis.split(",")[index]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else index
            self.handle = pynvml.nvmlDeviceGetHandleByIndex(phys)
            self.nvml = pynvml
        except Exception:
            self.nvml = None

    def _nvml_row(self):
        n, h = self.nvml, self.handle
        r = n.nvmlDeviceGetCurrentClocksEventReasons(h) if hasattr(n, "nvmlDeviceGetCurrentClocksEventReasons") else \
            n.nvmlDeviceGetCurrentClocksThrottleReasons(h)
        act = lambda bit: "Active" if (r & bit) else "Not Active"
        return [str(n.nvmlDeviceGetClockInfo(h, n.NVML_CLOCK_SM)), str(n.nvmlDeviceGetMaxClockInfo(h, n.NVML_CLOCK_SM)),
                "%.1f" % (n.nvmlDeviceGetPowerUsage(h) / 1e3), act(0x8), act(0x40), act(0x20), act(0x4)]

    def run(self):
        while not self.stop_flag:
            try:
                if self.nvml is not None:
                    self.rows.append(self._nvml_row())
                else:
                    o = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                       capture_output=True, text=True, timeout=5).stdout.strip()
                    if o:
                        self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.05 if self.nvml is not None else 0.1)

    def summary(self):
        sm = sorted(float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit())
        reasons = set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            for nme, v in zip(names, r[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nme)
        mx = max((float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()), default=None)
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": mx, "reasons": sorted(reasons), "samples": len(self.rows)}


def host_threads():
    return len(os.sched_getaffinity(0)) or 1


def cpu_reference_leg(a, steps, warmup):
    """The reference's padded / batched formulation of the step (oracle/baseline_batched.py; within 1.3x of the reference's
    own lib/sttran.py on the same video, tests/test_cpu_baseline.py) on the host cores, on a bounded sample of the workload."""
    from nlvsgg_b200 import shapes, synth
    from oracle import baseline, baseline_batched, cref
    torch.set_num_threads(host_threads())
    n = a.cpu_sample_videos
    entries = make_videos(a, 0, n, draw_fn=cref.draw_union_boxes, with_gt=True)
    tmpl = shapes.sttran_template() if a.arch == "sttran" else shapes.dsg_template()
    sd = synth.make_state_dict(tmpl, 0)
    if a.arch == "sttran":
        sec, frames = baseline_batched.time_cpu_steps(sd, entries, "sgdet", steps, warmup)
        how = "reference formulation (padded frames / windows, batched nn.MultiheadAttention)"
    else:
        sec, frames = baseline.time_cpu_steps(sd, entries, "sgdet", a.arch, steps, warmup)
        how = "segment-wise restatement"
    return {"value": frames / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
            "sample": f"{n} of the {a.videos} videos of one step ({frames} frames) per CPU step, {steps} step(s) after {warmup} warm-up: "
                      f"fwd + loss + bwd per video, clip + AdamW; torch CPU fp32; {how}; per-frame cost is independent of the batch size, "
                      f"so frames/s extrapolates to the full step",
            "same_config": False, "ms_per_step": sec * 1e3}


def run_reference(a):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    if a.config == "c5":
        return run_c5(a, reference=True)
    cpu = cpu_reference_leg(a, max(a.steps, 1), max(a.warmup, 1))
    val = cpu["value"]
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": cpu["ms_per_step"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic", "config": {"workload": workload_name(a), "sample": cpu["sample"]},
            "cpu_baseline": cpu, "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


PROFILE_STEPS = 5      # instrumented steps behind the roofline numbers (a single step's per-call times wander by a few percent)


def profile_one_step(step_fn):
    """Run `step_fn` (PROFILE_STEPS steps) with the sequencer's per-call CUDA-event timing on
    -> [(entry, ms, flops, bytes, m, n, k, dtype)] of all of them."""
    import ctypes
    from nlvsgg_b200 import _C
    lib = _C.lib()
    lib.nlv_profile_read.restype = ctypes.c_longlong
    torch.cuda.synchronize()
    lib.nlv_profile(1)
    step_fn()
    lib.nlv_profile(0)
    buf = ctypes.create_string_buffer(4 << 20)
    lib.nlv_profile_read(buf, ctypes.c_longlong(len(buf)))
    recs = []
    for line in buf.value.decode().splitlines():
        f = line.split("\t")
        recs.append((f[0], float(f[1]), float(f[2]), float(f[3]), int(f[4]), int(f[5]), int(f[6]), int(f[7])))
    return recs


def bind_to_gpu_cores(local: int):
    """Best effort: run this rank (and therefore first-touch its pinned staging buffers) on the cores NVML reports as local to
    its GPU."""
    try:
        import pynvml
        pynvml.nvmlInit()
        try:
            h = pynvml.nvmlDeviceGetHandleByUUID("GPU-" + str(torch.cuda.get_device_properties(local).uuid))
        except Exception:
            h = pynvml.nvmlDeviceGetHandleByIndex(local)
        ncpu = os.cpu_count() or 1
        words = pynvml.nvmlDeviceGetCpuAffinity(h, (ncpu + 63) // 64)
        near = {i for i in range(ncpu) if (int(words[i // 64]) >> (i % 64)) & 1}
        allowed = os.sched_getaffinity(0)
        pick = sorted(near & allowed)
        if pick and len(pick) < len(allowed):
            bind_to_gpu_cores.original = allowed          # restored before the CPU-baseline leg
            os.sched_setaffinity(0, pick)
            return f"bound to {len(pick)} GPU-local cores of {len(allowed)}"
        return f"all {len(allowed)} visible cores are GPU-local" if pick else "no GPU-local core visible"
    except Exception as e:  # NVML missing / not permitted: run unbound
        return f"unbound ({type(e).__name__})"


def peaks():
    try:
        return json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        return {}


def rooflines(recs, step_ms, n_stored_values, adamw_ms, n_params):
    """GEMM (tensor) roofline + HBM fractions of the gather / attention / normalisation / optimiser kernel classes from the
    per-call records of one instrumented step.  achieved = algorithmic FLOPs (bytes) / CUDA-event time."""
    pk = peaks()
    tpeak = pk.get("bf16_tflops_sustained") or 1400.0
    hpeak = pk.get("hbm_gbs") or 6650.0
    src = "MEASURED_PEAKS.json" if pk else "fallback (B200_PROFILING.md)"
    reps = PROFILE_STEPS
    recs = [(r[0], r[1] / reps, r[2] / reps, r[3] / reps) + tuple(r[4:]) for r in recs]     # per-step shares of every record
    g = [r for r in recs if r[0] in ("nlv_gemm", "nlv_conv3x3_dgrad") and r[7] == 1 and r[2] > 0]      # every launch of the tcgen05 kernel
    roof = None
    if g:
        tsum, fsum = sum(r[1] for r in g) / 1e3, sum(r[2] for r in g)
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "gemm_traffic.json"))).get("dram_bytes_per_launch_avg")
        except Exception:
            pass
        ach = fsum / tsum / 1e12
        roof = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05)", "achieved": ach, "peak": tpeak, "unit": "TFLOP/s",
                "frac": ach / tpeak, "traffic": traffic, "launches_per_step": len(g) // reps, "algorithmic_flop_per_step": fsum,
                "algorithmic_flop_per_launch_avg": fsum * reps / len(g), "avg_launch_ms": tsum * 1e3 * reps / len(g),
                "kernel_share_of_step": tsum * 1e3 / step_ms, "peak_source": f"{src} bf16_tflops_sustained (kernel timed inside a long step)"}

    def hbm(names, extra_bytes=0.0):
        rs = [r for r in recs if r[0] in names and r[3] > 0]
        if not rs:
            return None
        ms, by = sum(r[1] for r in rs), sum(r[3] for r in rs) + extra_bytes
        return {"bound": "hbm", "achieved": by / ms / 1e6, "peak": hpeak, "unit": "GB/s", "frac": by / ms / 1e6 / hpeak, "launches": len(rs) // reps,
                "ms_per_step": ms, "algorithmic_bytes_per_step": by}
    extra = {
        "gather_union_rows": hbm(("nlv_union_unpack", "nlv_union_unpack12", "nlv_nchw_to_rows"), float(n_stored_values)),
        "attention_fwd": hbm(("nlv_attn_fwd", "nlv_attn_fwd_drop", "nlv_attn_fwd_padkeys")),
        "attention_bwd": hbm(("nlv_attn_bwd", "nlv_attn_bwd_drop", "nlv_attn_bwd_sorted")),
        "layernorm": hbm(("nlv_layernorm_fwd", "nlv_layernorm_bwd", "nlv_layernorm_bwd_drop", "nlv_layernorm_bwd_fused")),
        "batchnorm": hbm(("nlv_bn_stats", "nlv_bn_apply", "nlv_bn_bwd", "nlv_bn_bwd_colsum")),
        "mask_conv1": hbm(("nlv_mask_conv1_fwd", "nlv_bn_apply_maxpool", "nlv_mask_conv1_dw")),
        "bias_grad_colsum": hbm(("nlv_colsum",)),
    }
    if adamw_ms:
        by = 30.0 * n_params     # p, g, m, v read (16 B); p, m, v + bf16 mirror written (14 B)
        extra["adamw"] = {"bound": "hbm", "achieved": by / adamw_ms / 1e6, "peak": hpeak, "unit": "GB/s", "frac": by / adamw_ms / 1e6 / hpeak,
                          "launches": 1, "ms_per_step": adamw_ms, "algorithmic_bytes_per_step": by}
    extra["peak_source"] = f"{src} hbm_gbs"
    by_entry = {}
    for r in recs:
        c = by_entry.setdefault(r[0], [0, 0.0])
        c[0] += 1
        c[1] += r[1]
    return roof, extra, {k: {"calls": v[0] // reps, "ms": round(v[1], 4)} for k, v in sorted(by_entry.items(), key=lambda kv: -kv[1][1])}


def dropin_legs(dev, precision):
    """The path tools/train_STTran.py / tools/test_STTran.py drive: one video per call through lib.sttran.STTran.
    C1 = predcls inference (20 frames, ~6 boxes); plus one sgdet training iteration (forward, the script's own torch losses,
    backward through the module, clip_grad_norm_, AdamW)."""
    import copy
    from nlvsgg_b200 import shapes, synth
    from nlvsgg_b200.lib.sttran import STTran
    out = {}
    for mode, train in (("predcls", False), ("sgdet", True)):
        e, _ = synth.synth_video(0, 20, 6, mode)
        e = {k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in e.items()}
        m = STTran(mode, 3, 6, 17, synth.AG_OBJECT_CLASSES, 1, 3, "wk", True, 2048, precision=precision)
        m.load_state_dict(synth.make_state_dict(shapes.sttran_template(), 0))
        m = m.to(dev)
        frames = int(e["im_idx"].max().item()) + 1
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        if not train:
            m.eval()

            def it():
                with torch.no_grad():
                    return m(dict(e))["attention_distribution"]
        else:
            m.train()
            opt = torch.optim.AdamW(m.parameters(), lr=1e-5)
            ce, bce = torch.nn.CrossEntropyLoss(), torch.nn.BCELoss()
            att = torch.tensor([int(x[0]) for x in e["attention_gt"]], device=dev)
            R = len(e["spatial_gt"])
            spa, con = torch.zeros(R, 6, device=dev), torch.zeros(R, 17, device=dev)
            for i in range(R):
                spa[i, e["spatial_gt"][i]] = 1.0
                con[i, e["contacting_gt"][i]] = 1.0

            def it():
                pred = m(dict(e))
                loss = (ce(pred["distribution"], pred["labels"]) + ce(pred["attention_distribution"], att)
                        + bce(pred["spatial_distribution"], spa) + bce(pred["contacting_distribution"], con))
                opt.zero_grad()
                loss.backward()
                torch.nn.utils.clip_grad_norm_(m.parameters(), max_norm=5, norm_type=2)
                opt.step()
                return loss
        for _ in range(3):
            it()
        torch.cuda.synchronize()
        n = 20
        t0 = time.perf_counter()
        ev0.record()
        for _ in range(n):
            r = it()
        ev1.record()
        float(r.flatten()[0].item())
        wall = (time.perf_counter() - t0) / n * 1e3
        ms = ev0.elapsed_time(ev1) / n
        key = "c1_predcls_inference" if not train else "sgdet_train_iteration"
        out[key] = {"frames": frames, "pairs": int(e["pair_idx"].shape[0]), "ms_per_call": ms, "host_wall_ms_per_call": wall,
                    "frames_per_s": frames / (ms / 1e3),
                    "path": "lib.sttran.STTran.forward(entry)" + (" + torch losses + backward + clip_grad_norm_ + torch.optim.AdamW" if train else " under no_grad")}
        del m
    return out


def parity_of_timed_mode(a, dev):
    """Max relative error of the timed precision mode's relation / object logits against the CPU oracle on a small video
    (training-mode forward, batch statistics) — the number the north star bounds at 1e-3 for fp32-accumulated logits."""
    from nlvsgg_b200 import engine as E, model as M, shapes, synth
    from oracle import cref, model as omodel
    entry, _ = synth.synth_video(11, 8, 6, "sgdet", draw_fn=cref.draw_union_boxes)
    tmpl = shapes.sttran_template() if a.arch == "sttran" else shapes.dsg_template()
    sd = synth.make_state_dict(tmpl, 0)
    fwd = omodel.sttran_forward if a.arch == "sttran" else omodel.dsg_forward
    with torch.no_grad():
        want = fwd({k: v.clone() for k, v in sd.items()}, entry, "sgdet", training=True)
    res = {}
    for prec in dict.fromkeys((a.precision, "bf16x3")):
        P = {k: v.to(dev) for k, v in sd.items()}
        batch, plan = M.make_batch([{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in entry.items()}], dev, "sgdet", dsg=a.arch == "dsg")
        f = M.sttran_forward if a.arch == "sttran" else M.dsg_forward
        out, _ = f(E.Kernels(prec), P, batch, plan, "sgdet", True, False)
        l26 = out["logits26"].float().cpu()
        errs = {"attention_logits": (l26[:, :3] - want["attention_distribution"]).abs().max().item() / want["attention_distribution"].abs().max().item(),
                "spatial_prob": (torch.sigmoid(l26[:, 3:9]) - want["spatial_distribution"]).abs().max().item(),
                "contacting_prob": (torch.sigmoid(l26[:, 9:]) - want["contacting_distribution"]).abs().max().item(),
                "object_logits": (out["distribution"].float().cpu() - want["distribution"]).abs().max().item() / want["distribution"].abs().max().item()}
        res[prec] = {k: float(f"{v:.3e}") for k, v in errs.items()}
    return {"vs": "oracle/model.py (CPU fp32) on an 8-frame synthetic video, training-mode forward", "max_rel_err": res,
            "north_star_bound": 1e-3}


def _evaluator(mode="sgdet"):
    from nlvsgg_b200 import synth
    from nlvsgg_b200.lib.evaluation_recall import SceneGraphEvaluator
    ev = SceneGraphEvaluator(mode=mode, AG_object_classes=synth.AG_OBJECT_CLASSES, AG_all_predicates=synth.AG_RELATIONS,
                             AG_attention_predicates=synth.AG_ATTENTION, AG_spatial_predicates=synth.AG_SPATIAL,
                             AG_contacting_predicates=synth.AG_CONTACTING, iou_threshold=0.5, constraint="with")
    ev.register_container()
    return ev


def _oracle_eval_rate(base, n=8):
    """frames/s of the numpy restatement of lib/evaluation_recall.py (oracle/evaluator.py) on `n` videos, one host thread."""
    from nlvsgg_b200 import synth
    from oracle import evaluator as oe
    ev = oe.Evaluator("sgdet", synth.AG_OBJECT_CLASSES, synth.AG_RELATIONS, synth.AG_ATTENTION, synth.AG_SPATIAL, synth.AG_CONTACTING,
                      iou_threshold=0.5, constraint="with")
    ev.register_container()
    t0 = time.perf_counter()
    fr = 0
    for pred, gt in base[:n]:
        ev.evaluate_scene_graph(gt, {k: (v.clone() if torch.is_tensor(v) else v) for k, v in pred.items()})
        fr += len(gt)
    sec = time.perf_counter() - t0
    return fr / sec, fr, sec


def run_c5(a, reference=False):
    """Recall@10/20/50 (with / no / semi constraint) over the test-split shape: 1737 videos, ~54k frames."""
    import numpy as np
    from nlvsgg_b200 import synth
    rng = np.random.default_rng(5)
    counts = np.clip(np.round(rng.gamma(shape=3.2, scale=9.8, size=a.videos)), 3, 121).astype(int)
    counts[0] = 121
    distinct = min(40, a.videos)       # generation is python-loop bound: 40 distinct videos tiled over the split
    base = [synth.synth_pred("sgdet", 5000 + i, int(counts[i]), 6, 0.05, saturate=(i % 7 == 0)) for i in range(distinct)]
    order = [i % distinct for i in range(a.videos)]
    n_frames = sum(len(base[j][1]) for j in order)
    metric = "recall_at_k_eval_frames_per_sec"
    wl = f"Recall@10/20/50 with / no / semi constraint, {a.videos} synthetic AG test videos ({n_frames} frames, ~6 predicted boxes, 3-5 GT objects)"
    rate, fr, csec = _oracle_eval_rate(base)
    cpu = {"value": rate, "unit": UNIT, "cores": 1, "kind": "port",
           "sample": f"8 videos / {fr} frames, numpy restatement of lib/evaluation_recall.py (oracle/evaluator.py), one host thread (the reference is single-threaded python)"}
    if reference:
        line = {"impl": "reference", "metric": metric, "value": rate, "unit": UNIT, "n_gpus": 1, "steps": 1, "warmup": 0, "ms_per_step": csec * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
                "config": {"workload": wl, "sample": cpu["sample"]}, "cpu_baseline": cpu,
                "e2e": {"value": rate, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
        print(json.dumps(line), flush=True)
        return
    from nlvsgg_b200 import _C
    _C.lib()
    dev = torch.device("cuda", int(os.environ.get("LOCAL_RANK", "0")))
    on_dev = [{k: (v.to(dev) if torch.is_tensor(v) else v) for k, v in base[j][0].items()} for j in range(distinct)]
    items = [(base[j][1], on_dev[j]) for j in order]            # predictions as the model leaves them (device), python GT annotations
    ev = _evaluator()
    t0 = time.perf_counter()
    ev.evaluate_videos([(g, dict(p)) for g, p in items])          # first pass: walks and caches the python GT annotations
    torch.cuda.synchronize()
    first_pass = time.perf_counter() - t0
    cache = ev._gt_cache
    times = []
    l0 = _C.launch_count()
    for _ in range(max(a.steps, 1)):
        ev = _evaluator()
        ev._gt_cache = cache                                      # ground truth is static across epochs
        batch = [(g, dict(p)) for g, p in items]
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        ev.evaluate_videos(batch)
        ev.calculate_mean_recall()
        torch.cuda.synchronize()
        times.append(time.perf_counter() - t0)
    launches = _C.launch_count() - l0
    sec = min(times)
    kern = getattr(ev, "last_kernel_ms", None)
    kb = getattr(ev, "last_algorithmic_bytes", 0)
    pk = peaks()
    hp = pk.get("hbm_gbs", 6650.0)
    line = {"metric": metric, "value": (n_frames / (kern / 1e3)) if kern else n_frames / sec, "unit": UNIT, "n_gpus": 1, "steps": a.steps,
            "warmup": 1, "ms_per_step": (kern if kern else sec * 1e3), "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f64+int", "data": "synthetic", "config": {"workload": wl, "config": "c5"},
            "e2e": {"value": n_frames / sec, "unit": UNIT, "ms_per_step": sec * 1e3, "h2d_bytes_per_step": getattr(ev, "last_h2d_bytes", None),
                    "d2h_bytes_per_step": getattr(ev, "last_d2h_bytes", None),
                    "first_epoch_ms": first_pass * 1e3,
                    "path": "SceneGraphEvaluator.evaluate_videos(device predictions + python GT) -> result_dict: GT arrays (cached per annotation after the "
                            "first epoch), one pinned upload, one kernel launch, one read-back, numpy booking, mean recall"},
            "gpu_launches": int(launches),
            "roofline": ({"bound": "hbm", "kernel": "recall_match_kernel", "achieved": kb / (kern / 1e3) / 1e9, "peak": hp, "unit": "GB/s",
                          "frac": kb / (kern / 1e3) / 1e9 / hp, "traffic": None,
                          "note": "latency-bound by construction: ~1-2 KB of input per frame (SURVEY 8d)"} if kern else None),
            "cpu_baseline": cpu,
            "recall": {f"R@{k}": float(np.mean(ev.result_dict["sgdet_recall"][k])) for k in (10, 20, 50)}}
    print(json.dumps(line), flush=True)


def main():
    a = parse()
    if a.impl == "reference":
        return run_reference(a)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — the hot path has no CPU fallback (use --impl reference for the CPU arm)")
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if a.config == "c5":
        if rank == 0:
            run_c5(a)
        return
    numa = bind_to_gpu_cores(local)
    # one process per GPU: keep the host-side torch ops of N ranks from oversubscribing the cores (the reference pins 4
    # threads itself, tools/train_STTran.py:40); the CPU-baseline leg raises this again for its own measurement
    torch.set_num_threads(max(1, min(4, host_threads() // max(world, 1))))
    if world > 1:
        import datetime
        torch.distributed.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    from nlvsgg_b200 import _C, featfile, model as M, ops, shapes, synth
    from nlvsgg_b200.trainer import Trainer
    _C.lib()

    if a.config == "c1":
        legs = dropin_legs(dev, a.precision)
        c1 = legs["c1_predcls_inference"]
        line = {"metric": "sttran_predcls_infer_frames_per_sec", "value": c1["frames_per_s"], "unit": UNIT, "n_gpus": 1, "steps": 20, "warmup": 3,
                "ms_per_step": c1["ms_per_call"], "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": a.precision,
                "data": "synthetic", "config": {"workload": "STTran PredCls inference, 1 synthetic AG video (20 frames, ~6 boxes/frame) through lib.sttran.STTran"},
                "dropin": legs}
        print(json.dumps(line), flush=True)
        return

    tmpl = shapes.sttran_template() if a.arch == "sttran" else shapes.dsg_template()
    sd = synth.make_state_dict(tmpl, 0)
    # dropout as the reference trains (p = 0.1 on attention weights, residual branches, FFN: lib/transformer.py:7-29,36-57); the
    # fp32-grade parity modes run without it (masks cannot be matched to torch's RNG, parity is defined at p = 0)
    p_drop = 0.1 if a.precision == "bf16" else 0.0
    trainer = Trainer({k: v.to(dev) for k, v in sd.items()}, "sgdet", a.arch, a.precision, device=dev, dropout=p_drop)
    entries = make_videos(a, rank, a.videos, with_gt=True)
    tmpdir = None
    if a.input == "packed":
        # the loader path: one packed file per video (written once here, outside every timed region), read straight into pinned staging
        base = "/dev/shm" if os.path.isdir("/dev/shm") else None
        tmpdir = tempfile.mkdtemp(prefix=f"nlv_bench_r{rank}_", dir=base)
        paths = featfile.write_videos(tmpdir, entries)
        host = featfile.Loader(pin=True, depth=1).load(paths)
        density = float(sum(featfile.read_header(p)["union_nnz"] for p in paths)) / max(1, sum(host.n_pairs) * 49 * 2048)
        # bytes of the stored union values (what the unpack kernel reads besides bitmap + offsets)
        n_stored = {2: 2 * int(host.union_feat.numel()), 3: int(host.union_feat.numel()) + int(host.union_hx.numel()) if host.union_rows == 3 else 0
                    }.get(host.union_rows, 0)
        enc = {2: "zero-suppressed (density %.3f), 16-bit values" % density,
               3: "zero-suppressed (density %.3f), 12-bit values (low byte + 4-bit high-byte code over a per-row base; lossless)" % density}
        in_fmt = (f"packed per-video feature files (featfile.py): bf16 features, channels-last union rows, "
                  f"{enc.get(host.union_rows, 'dense')}, create_dis pairs; decoded inside the step")
    else:
        host = M.collate(entries, "sgdet", pin=True)
        n_stored = 0
        in_fmt = "fp32 NCHW entry tensors (the reference's entry contract)"
    frames = sum(int(f.max()) + 1 for f in host.frame_ids if len(f))
    h2d_bytes = M.input_bytes(host)
    resident = M.upload(host, dev, rasterise=False)      # the mask rasteriser (a3) runs inside every step

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def resident_step():
        b = M.Batch()
        b.__dict__.update(resident.__dict__)
        return trainer.step(b)

    for _ in range(a.warmup):
        resident_step()
    barrier()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = _C.launch_count()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(a.steps):
        loss = resident_step()
    ev1.record()
    barrier()
    launches = _C.launch_count() - l0
    ms = ev0.elapsed_time(ev1) / a.steps
    t = torch.tensor([ms], device=dev)
    if world > 1:
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
    ms = float(t.item())
    fr = torch.tensor([float(frames)], device=dev)
    if world > 1:
        torch.distributed.all_reduce(fr)
    total_frames = float(fr.item())
    value = total_frames / (ms / 1e3)
    loss_v = float(loss.item())

    def e2e_loop(hb):
        """every step copies its own inputs from pinned host memory; the copy for step i+1 is issued on a side stream before step i
        computes, so transfers and compute overlap in steady state; every step's loss is copied back to the host behind the step
        and read there one step later (while the next step runs), the last one before the closing event"""
        import gc
        trainer.step_from_host(hb).item()
        depth = int(os.environ.get("NLV_BENCH_PREFETCH", "3"))
        lag = int(os.environ.get("NLV_BENCH_LOSS_LAG", "2"))
        t_step, t_pref, t_wait = [], [], []      # host milliseconds per step: enqueue of the step, of the next copies, wait for a loss

        def run(n_steps):
            """n_steps of the pipelined loop: copies up to `depth` steps ahead, every loss read `lag` steps late"""
            queue = [trainer.prefetch(hb) for _ in range(min(depth, n_steps))]
            tickets, last = [], None
            for i in range(n_steps):
                h0 = time.perf_counter()
                trainer.step_pipelined(queue.pop(0), None)                        # step i is enqueued first ...
                h1 = time.perf_counter()
                if i + depth < n_steps:
                    queue.append(trainer.prefetch(hb))                            # ... then the copies of step i+depth (behind the earlier ones)
                h2 = time.perf_counter()
                tickets.append(trainer.last_ticket)
                if len(tickets) > lag:
                    last = trainer.loss_value(tickets.pop(0))    # step i-lag's loss (its own D2H copy), read while later steps run
                h3 = time.perf_counter()
                t_step.append(1e3 * (h1 - h0)); t_pref.append(1e3 * (h2 - h1)); t_wait.append(1e3 * (h3 - h2))
            for tk in tickets:
                last = trainer.loss_value(tk)
            return last

        # warm-up THROUGH the pipelined path (the copy stream, the device blocks of three batches in flight and the staging ring are
        # created on first use: runs that paid for them inside the timed region showed 0.3-1.3 s of one-time stall, read as a
        # 2-3x slower step)
        run(max(3, a.warmup))
        torch.cuda.synchronize()
        del t_step[:], t_pref[:], t_wait[:]
        # the host is on this loop's critical path (it reads a loss every step): no cyclic-GC pauses inside it
        gc.collect(); gc.freeze(); gc.disable()
        barrier()
        ev0.record()
        lv = run(a.steps)
        ev1.record()
        gc.enable(); gc.unfreeze()
        med = lambda v: sorted(v)[len(v) // 2]
        e2e_loop.host = {"enqueue_step_ms_median": round(med(t_step), 3), "enqueue_copies_ms_median": round(med(t_pref), 3),
                         "wait_loss_ms_median": round(med(t_wait), 3), "enqueue_step_ms_max": round(max(t_step), 3),
                         "enqueue_copies_ms_max": round(max(t_pref), 3), "wait_loss_ms_max": round(max(t_wait), 3)}
        barrier()
        t = torch.tensor([ev0.elapsed_time(ev1) / a.steps], device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item()), lv

    def h2d_alone(hb, reps=4):
        """the step's host->device copies alone (every rank at once): the link rate the e2e step is measured against"""
        warm = [M.upload(hb, dev, rasterise=False) for _ in range(reps)]      # device blocks come from the caching allocator afterwards
        torch.cuda.synchronize()
        del warm
        barrier()
        torch.cuda.synchronize()
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        keep = [M.upload(hb, dev, rasterise=False) for _ in range(reps)]
        c1.record()
        torch.cuda.synchronize()
        del keep
        t = torch.tensor([c0.elapsed_time(c1) / reps], device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        return float(t.item())

    # the clock sampler covers the timed region above; it is stopped here: the end-to-end loop has the host on its critical path and
    # the sampler's NVML queries were one suspect for runs in which every step of that loop took 2-3x longer
    sampler.stop_flag = True
    sampler.join(timeout=2)

    e2e = None
    if not a.no_e2e:
        ems, lv = e2e_loop(host)
        copy_ms = h2d_alone(host)
        e2e = {"value": total_frames / (ems / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
               "ms_per_step": ems, "last_loss": lv, "input_pipelining": "H2D copies run up to three steps ahead of the compute on a side stream; each loss is read two steps late from its own pinned copy",
               "host_format": in_fmt, "cpu_binding": numa, "host_side": getattr(e2e_loop, "host", None),
               "h2d_alone_ms_per_step": copy_ms, "h2d_alone_gb_per_s_per_gpu": h2d_bytes / copy_ms / 1e6,
               "h2d_gb_per_s_if_copy_bound": h2d_bytes / ems / 1e6}
    # ---- rooflines: one instrumented step (every rank runs it: it contains the gradient allreduce) ----
    step_s, step_e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    recs = []

    def instrumented():
        step_s.record()
        for _ in range(PROFILE_STEPS):
            resident_step()
        step_e.record()
    if rank == 0:
        recs = profile_one_step(instrumented)
    else:
        instrumented()
    torch.cuda.synchronize()
    roof = extra = by_entry = None
    if rank == 0 and recs:
        # the optimiser kernel alone (events around optimizer_step of an extra step's gradients)
        o0, o1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        adamw_ms = None
        if world == 1:
            torch.cuda.synchronize()
            o0.record()
            trainer.optimizer_step()
            o1.record()
            torch.cuda.synchronize()
            adamw_ms = o0.elapsed_time(o1)
        roof, extra, by_entry = rooflines(recs, step_s.elapsed_time(step_e) / PROFILE_STEPS, n_stored, adamw_ms, trainer.n_params)

    if world > 1:
        torch.distributed.barrier()
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        if tmpdir:
            shutil.rmtree(tmpdir, ignore_errors=True)
        return
    extras = {}
    if world == 1 and not a.no_extras:
        # (1) the same step fed with the reference's fp32 NCHW entry tensors (4.9 GB per step over PCIe)
        if a.input == "packed" and not a.no_e2e and a.config == "c2":
            host32 = M.collate(entries, "sgdet", pin=True)
            ems32, _ = e2e_loop(host32)
            res32 = M.upload(host32, dev, rasterise=False)
            for _ in range(2):
                trainer.step(res32)
            torch.cuda.synchronize()
            ev0.record()
            for _ in range(a.steps):
                trainer.step(res32)
            ev1.record()
            torch.cuda.synchronize()
            extras["fp32_entry_contract_input"] = {"resident_ms_per_step": ev0.elapsed_time(ev1) / a.steps,
                                                   "resident_frames_per_s": total_frames / (ev0.elapsed_time(ev1) / a.steps / 1e3),
                                                   "e2e_ms_per_step": ems32, "e2e_frames_per_s": total_frames / (ems32 / 1e3),
                                                   "h2d_bytes_per_step": M.input_bytes(host32)}
            del host32, res32
        # (2) the parity mode (fp32-grade GEMMs through the three-term bf16 split, fp32 attention): same workload
        if a.precision == "bf16" and a.config == "c2":
            del trainer
            torch.cuda.empty_cache()
            tr3 = Trainer({k: v.to(dev) for k, v in sd.items()}, "sgdet", a.arch, "bf16x3", device=dev)
            for _ in range(2):
                tr3.step(resident)
            torch.cuda.synchronize()
            ev0.record()
            for _ in range(3):
                tr3.step(resident)
            ev1.record()
            torch.cuda.synchronize()
            extras["parity_mode_bf16x3"] = {"ms_per_step": ev0.elapsed_time(ev1) / 3, "frames_per_s": total_frames / (ev0.elapsed_time(ev1) / 3 / 1e3),
                                            "note": "logits within 1e-3 of the fp32 reference (see `parity`)"}
            del tr3
            torch.cuda.empty_cache()
        # (3) the single-video drop-in path
        extras["dropin"] = dropin_legs(dev, a.precision)
    cpu = parity = None
    if world == 1 and not a.no_cpu_baseline:
        if getattr(bind_to_gpu_cores, "original", None):   # the CPU baseline may use every core the process was given
            os.sched_setaffinity(0, bind_to_gpu_cores.original)
        parity = parity_of_timed_mode(a, dev)
        cpu = cpu_reference_leg(a, 1, 1)
    if tmpdir:
        shutil.rmtree(tmpdir, ignore_errors=True)
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "bf16x3": "bf16x3", "fp32": "f32"}[a.precision], "data": "synthetic",
            "config": {"workload": workload_name(a), "config": a.config, "videos_per_gpu": a.videos, "frames_per_step_all_gpus": total_frames,
                       "pairs_per_gpu": int(sum(host.n_pairs)), "boxes_per_gpu": int(sum(host.n_boxes)),
                       "parallelism": f"dp{world}", "resident_input_format": in_fmt, "dropout_p": p_drop,
                       "l2": "per-step working set (>10 GB of activations) exceeds the 126 MB L2; no flush needed",
                       "step": "input decode + mask rasterise + forward + fused losses + backward + allreduce + clip + AdamW"},
            "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "roofline_extra": extra,
            "cpu_baseline": cpu, "parity": parity, "loss": loss_v, "step_entries": by_entry}
    line.update(extras)
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
