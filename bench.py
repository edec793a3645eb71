#!/usr/bin/env python
"""Benchmark of the hot path (BASELINE.json): STTran SGDet training step — forward, losses, backward, gradient
allreduce, clip + AdamW — over a batch of synthetic Action-Genome-shaped videos, frames/s.

    python bench.py --gpus N --steps K --warmup W            # this repo (sm_100a kernels)
    python bench.py --impl reference --steps K --warmup W     # the reference algorithm on the host CPU cores

One JSON line on stdout (rank 0).  See DESIGN.md §Measurement for every field.
"""
import argparse
import json
import os
import subprocess
import sys
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
    ap.add_argument("--videos", type=int, default=64, help="videos per GPU per step (BASELINE config C2: 64)")
    ap.add_argument("--frames", type=int, default=30, help="mean frames per video (U{f-10..f+10})")
    ap.add_argument("--boxes", type=int, default=7, help="mean boxes per frame")
    ap.add_argument("--precision", default="bf16", choices=["bf16", "bf16x3", "fp32"])
    ap.add_argument("--arch", default="sttran", choices=["sttran", "dsg"])
    ap.add_argument("--cpu-sample-videos", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-packed-e2e", action="store_true", help="skip the extra e2e leg fed from bf16 feature buffers")
    return ap.parse_args()


def workload_name(a):
    return (f"{'STTran' if a.arch == 'sttran' else 'DSG-DETR'} SGDet training step, {a.videos} synthetic AG videos/GPU "
            f"(~{a.frames} frames, ~{a.boxes} VinVL 2048-d boxes/frame), {a.precision}")


def make_videos(a, rank, n, draw_fn=None):
    from nlvsgg_b200 import synth
    g = torch.Generator().manual_seed(777 + rank)
    out = []
    for i in range(n):
        frames = int(torch.randint(max(2, a.frames - 10), a.frames + 11, (1,), generator=g))
        e, _ = synth.synth_video(100000 * rank + i, frames, a.boxes, "sgdet", draw_fn=draw_fn, with_gt=False)
        out.append(e)
    return out


class ClockSampler(threading.Thread):
    """nvidia-smi clocks / throttle reasons sampled during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                   capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.rows.append([x.strip() for x in o.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

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


def run_reference(a):
    """The reference algorithm (oracle port of lib/sttran.py + tools/train_STTran.py step) on the host CPU cores."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from nlvsgg_b200 import shapes, synth
    from oracle import baseline, cref
    torch.set_num_threads(len(os.sched_getaffinity(0)) or 1)    # every core this process may run on
    n = a.cpu_sample_videos
    entries = make_videos(a, 0, n, draw_fn=cref.draw_union_boxes)
    tmpl = shapes.sttran_template() if a.arch == "sttran" else shapes.dsg_template()
    sd = synth.make_state_dict(tmpl, 0)
    sec, frames = baseline.time_cpu_steps(sd, entries, "sgdet", a.arch, a.steps, a.warmup)
    val = frames / sec
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": a.gpus, "steps": a.steps,
            "warmup": a.warmup, "ms_per_step": sec * 1e3, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": "f32", "data": "synthetic",
            "config": {"workload": workload_name(a), "sample": f"{n} of the {a.videos} videos of one step per CPU step"},
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
                             "sample": f"{n} videos / {frames} frames per step, fwd+loss+bwd+clip+AdamW, torch CPU fp32"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def bind_to_gpu_cores(local: int):
    """Best effort: run this rank (and therefore first-touch its pinned staging buffers) on the cores NVML reports as local to
    its GPU.  The e2e leg moves 4.9 GB per step over PCIe; pinned memory on the far NUMA node halves that bandwidth."""
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
    numa = bind_to_gpu_cores(local)
    # one process per GPU: keep the host-side torch ops of N ranks from oversubscribing the cores (the reference pins 4
    # threads itself, tools/train_STTran.py:40); the CPU-baseline leg raises this again for its own measurement
    torch.set_num_threads(max(1, min(4, len(os.sched_getaffinity(0)) // max(world, 1))))
    if world > 1:
        import datetime
        torch.distributed.init_process_group("nccl", device_id=dev, timeout=datetime.timedelta(seconds=180))
    from nlvsgg_b200 import _C, model as M, ops, shapes, synth
    from nlvsgg_b200.trainer import Trainer
    _C.lib()

    tmpl = shapes.sttran_template() if a.arch == "sttran" else shapes.dsg_template()
    sd = synth.make_state_dict(tmpl, 0)
    trainer = Trainer({k: v.to(dev) for k, v in sd.items()}, "sgdet", a.arch, a.precision, device=dev)
    entries = make_videos(a, rank, a.videos)
    host = M.collate(entries, "sgdet", pin=True)
    frames = sum(int(f.max()) + 1 for f in host.frame_ids if len(f))
    h2d_bytes = M.input_bytes(host)
    resident = M.upload(host, dev)
    resident.spatial_masks = None  # the mask rasteriser (a3) runs inside every step
    del entries

    def barrier():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def resident_step():
        b = M.Batch()
        b.__dict__.update(resident.__dict__)
        return trainer.step(M.ensure_masks(b))

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

    # ---- end to end: pinned host buffers -> device inside the step, loss read back every step ----
    e2e = None
    if not a.no_e2e:
        trainer.step_from_host(host).item()
        barrier()
        # every step copies its own inputs from pinned host memory; the copy for step i+1 is issued on a side
        # stream before step i computes, so transfers and compute overlap in steady state
        ev0.record()
        nxt = trainer.prefetch(host)
        for i in range(a.steps):
            loss_t, nxt = trainer.step_pipelined(nxt, host if i + 1 < a.steps else None)
            lv = loss_t.item()
        ev1.record()
        barrier()
        ems = ev0.elapsed_time(ev1) / a.steps
        t = torch.tensor([ems], device=dev)
        if world > 1:
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        e2e = {"value": total_frames / (float(t.item()) / 1e3), "unit": UNIT, "h2d_bytes_per_step": h2d_bytes, "d2h_bytes_per_step": 4,
               "ms_per_step": float(t.item()), "last_loss": lv, "input_pipelining": "H2D of step i+1 overlaps compute of step i (side stream)",
               "host_feature_dtype": "fp32 (the reference's entry contract)", "cpu_binding": numa}
        if a.precision == "bf16" and not a.no_packed_e2e and world == 1:   # informational leg: single-GPU runs only
            # Extra, NOT the headline: the same loop fed from the packed bf16 feature format.  In the bf16 compute
            # mode the step is bit-identical (the round-to-nearest moves from the device into the loader); the copy is half as long.
            host16 = M.repack(host, torch.bfloat16)
            trainer.step_from_host(host16).item()
            barrier()
            ev0.record()
            nxt = trainer.prefetch(host16)
            for i in range(a.steps):
                loss_t, nxt = trainer.step_pipelined(nxt, host16 if i + 1 < a.steps else None)
                lv16 = loss_t.item()
            ev1.record()
            barrier()
            t = torch.tensor([ev0.elapsed_time(ev1) / a.steps], device=dev)
            if world > 1:
                torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            e2e["packed_bf16_features"] = {"value": total_frames / (float(t.item()) / 1e3), "unit": UNIT, "ms_per_step": float(t.item()),
                                           "h2d_bytes_per_step": M.input_bytes(host16), "last_loss": lv16}
            del host16
    sampler.stop_flag = True
    sampler.join(timeout=2)

    # ---- roofline of the dominant kernel (the tcgen05 GEMM): events around every launch of one step ----
    # every rank runs this extra step (it contains the gradient allreduce); only rank 0 instruments it
    roof = None
    recs = []
    orig = ops.gemm

    def timed_gemm(a_, b_, out, **kw):
        if a_.dtype == torch.bfloat16 and not kw.get("force_simt"):
            s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            s.record()
            r = orig(a_, b_, out, **kw)
            e.record()
            m_, n_ = out.shape
            k_ = a_.shape[1] if kw.get("a_major", 0) == 0 else a_.shape[0]
            recs.append((s, e, 2.0 * m_ * n_ * k_))
            return r
        return orig(a_, b_, out, **kw)
    if rank == 0:
        ops.gemm = timed_gemm
    step_s, step_e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    step_s.record()
    resident_step()
    step_e.record()
    torch.cuda.synchronize()
    ops.gemm = orig
    if rank == 0 and recs:
        tsum = sum(s.elapsed_time(e) for s, e, _ in recs) / 1e3
        fsum = sum(f for _, _, f in recs)
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = peaks.get("bf16_tflops_sustained") or 1400.0
        ach = fsum / tsum / 1e12
        traffic = None
        try:
            traffic = json.load(open(os.path.join(ROOT, "profiles", "gemm_traffic.json"))).get("dram_bytes_per_launch_avg")
        except Exception:
            pass
        roof = {"bound": "tensor", "kernel": "gemm_tc_kernel (tcgen05)", "achieved": ach, "peak": peak, "unit": "TFLOP/s",
                "frac": ach / peak, "traffic": traffic, "launches_per_step": len(recs), "algorithmic_flop_per_step": fsum,
                "algorithmic_flop_per_launch_avg": fsum / len(recs), "avg_launch_ms": tsum * 1e3 / len(recs),
                "kernel_share_of_step": tsum * 1e3 / step_s.elapsed_time(step_e),
                "peak_source": "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback 1400 (B200_PROFILING.md sustained)"}

    if world > 1:
        torch.distributed.barrier()
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    cpu = None
    if world == 1 and not a.no_cpu_baseline:
        from oracle import baseline, cref
        if getattr(bind_to_gpu_cores, "original", None):   # the CPU baseline may use every core the process was given
            os.sched_setaffinity(0, bind_to_gpu_cores.original)
        torch.set_num_threads(len(os.sched_getaffinity(0)) or 1)
        n = a.cpu_sample_videos
        ce = make_videos(a, 0, n, draw_fn=cref.draw_union_boxes)
        sec, cfr = baseline.time_cpu_steps(synth.make_state_dict(tmpl, 0), ce, "sgdet", a.arch, 1, 1)
        cpu = {"value": cfr / sec, "unit": UNIT, "cores": torch.get_num_threads(), "kind": "port",
               "sample": f"{n} videos / {cfr} frames, 1 step after 1 warm-up, fwd+loss+bwd+clip+AdamW, torch CPU fp32"}
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": a.steps, "warmup": a.warmup,
            "ms_per_step": ms, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
            "dtype": {"bf16": "bf16", "bf16x3": "bf16x3", "fp32": "f32"}[a.precision], "data": "synthetic",
            "config": {"workload": workload_name(a), "videos_per_gpu": a.videos, "frames_per_step_all_gpus": total_frames,
                       "pairs_per_gpu": int(sum(host.n_pairs)), "boxes_per_gpu": int(sum(host.n_boxes)),
                       "parallelism": f"dp{world}", "l2": "per-step working set (>10 GB of activations) exceeds the 126 MB L2; no flush needed",
                       "step": "mask rasterise + forward + fused losses + backward + allreduce + clip + AdamW"},
            "clocks": sampler.summary(), "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "cpu_baseline": cpu,
            "loss": float(loss.item())}
    print(json.dumps(line), flush=True)
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
