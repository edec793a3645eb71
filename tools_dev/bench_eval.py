"""Dev/measurement: Recall@K evaluator throughput at BASELINE config C5 shape (many videos, ~31 frames each) —
CUDA kernel (one launch for all videos) vs the numpy oracle on a bounded sample.  Writes a JSON line."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np, torch
from nlvsgg_b200 import synth
from nlvsgg_b200.lib.evaluation_recall import SceneGraphEvaluator
from oracle.make_golden_eval import synth_pred
from oracle import evaluator as oe

n_videos = int(os.environ.get("VIDEOS", 400))
g = torch.Generator().manual_seed(5)
vids = []
for i in range(n_videos):
    frames = int(torch.randint(10, 60, (1,), generator=g))
    vids.append(synth_pred("sgdet", 9000 + i, frames, 6, 0.05, False))
n_frames = sum(len(gt) for _, gt in vids)
def mk():
    ev = SceneGraphEvaluator("sgdet", synth.AG_OBJECT_CLASSES, synth.AG_RELATIONS, synth.AG_ATTENTION, synth.AG_SPATIAL,
                             synth.AG_CONTACTING, 0.5, "with")
    ev.register_container(); return ev
cuda_vids = [(gt, {k: (v.cuda() if torch.is_tensor(v) else v) for k, v in pred.items()}) for pred, gt in vids]
def run():
    ev = mk()
    ev.evaluate_videos([(gt, dict(p)) for gt, p in cuda_vids])
    return ev
run(); torch.cuda.synchronize()
t0 = time.perf_counter(); ev = run(); torch.cuda.synchronize(); t_all = time.perf_counter() - t0
# kernel-only time
from nlvsgg_b200 import ops
import nlvsgg_b200.lib.evaluation_recall as ER
orig = ER.recall_match
times = {}
def timed(*a, **k):
    s, e = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    s.record(); r = orig(*a, **k); e.record(); torch.cuda.synchronize(); times["ms"] = s.elapsed_time(e); return r
ER.recall_match = timed; run(); ER.recall_match = orig
# oracle sample
sample = vids[:12]
o = oe.Evaluator("sgdet", synth.AG_OBJECT_CLASSES, synth.AG_RELATIONS, synth.AG_ATTENTION, synth.AG_SPATIAL, synth.AG_CONTACTING)
o.register_container()
t0 = time.perf_counter()
for pred, gt in sample:
    p = dict(pred); p["attention_distribution"] = torch.softmax(p["attention_distribution"], 1)
    o.evaluate_scene_graph(gt, p)
t_cpu = time.perf_counter() - t0
cpu_frames = sum(len(gt) for _, gt in sample)
print(json.dumps({"workload": f"Recall@K over {n_videos} synthetic videos / {n_frames} frames (C5 shape), sgdet", "frames": n_frames,
                  "cuda_end_to_end_s": t_all, "cuda_frames_per_s": n_frames / t_all,
                  "kernel_plus_copies_ms": times["ms"], "kernel_frames_per_s": n_frames / (times["ms"] / 1e3),
                  "cpu_oracle_frames_per_s": cpu_frames / t_cpu, "cpu_sample_frames": cpu_frames,
                  "R@20": float(np.mean(ev.result_dict["sgdet_recall"][20]))}))
