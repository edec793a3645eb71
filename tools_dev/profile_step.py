"""Dev: per-entry-point CUDA-event breakdown of one training step + host-side timings."""
import sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from nlvsgg_b200 import _C, model as M, ops, shapes, synth
from nlvsgg_b200.trainer import Trainer

class A: pass
a = A(); a.videos = int(os.environ.get("VIDEOS", 64)); a.frames = 30; a.boxes = 7; a.arch = os.environ.get("ARCH", "sttran")
a.precision = os.environ.get("PREC", "bf16")
dev = torch.device("cuda")
tmpl = shapes.sttran_template() if a.arch == "sttran" else shapes.dsg_template()
tr = Trainer({k: v.to(dev) for k, v in synth.make_state_dict(tmpl, 0).items()}, "sgdet", a.arch, a.precision, device=dev)
host = M.collate(bench.make_videos(a, 0, a.videos), "sgdet", pin=True)
res = M.upload(host, dev)
def step():
    b = M.Batch(); b.__dict__.update(res.__dict__)
    b.spatial_masks = ops.union_mask_pairs(b.boxes, b.pair_idx, 27, -0.5)
    return tr.step(b)
for _ in range(2): step()
torch.cuda.synchronize()
# host-side cost of one step (no sync inside)
t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
print(f"host enqueue time {1e3*(t1-t0):.1f} ms, step wall {1e3*(t2-t0):.1f} ms")
t0 = time.perf_counter(); plan = M.make_plan(res, dev, "sgdet", a.arch == "dsg"); t1 = time.perf_counter(); lab = M.make_labels(res, dev, "sgdet"); t2 = time.perf_counter()
print(f"make_plan {1e3*(t1-t0):.1f} ms, make_labels {1e3*(t2-t1):.1f} ms")
ops.PROFILE = {}
step()
summ = ops.profile_summary()
ops.PROFILE = None
tot = sum(v[1] for v in summ.values())
print(f"sum of per-call event times: {tot:.2f} ms over {sum(v[0] for v in summ.values())} calls")
agg = {}
for k, (n, ms) in summ.items():
    key = k if not k.startswith("nlv_gemm[") else k.split(" m=")[0] + "]"
    c = agg.setdefault(key, [0, 0.0]); c[0] += n; c[1] += ms
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:8.3f} ms {100*ms/tot:5.1f}%  n={n:4d}  {k}")
print("--- top individual GEMMs")
for k, (n, ms) in sorted(((k, v) for k, v in summ.items() if k.startswith("nlv_gemm[")), key=lambda kv: -kv[1][1])[:25]:
    import re
    m_, n_, k_ = map(int, re.findall(r"[mnk]=(\d+)", k))
    print(f"{ms:8.3f} ms n={n:3d} {k}  {2*m_*n_*k_*n/ms/1e9:8.1f} TFLOP/s")
