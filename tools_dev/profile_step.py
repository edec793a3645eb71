"""Dev: per-entry-point CUDA-event breakdown of one training step (sequencer profile) + host-side timings."""
import ctypes, sys, os, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from nlvsgg_b200 import _C, model as M, ops, shapes, synth
from nlvsgg_b200.trainer import Trainer

class A: pass
a = A(); a.videos = int(os.environ.get("VIDEOS", 64)); a.frames = 30; a.boxes = 7; a.arch = os.environ.get("ARCH", "sttran")
a.precision = os.environ.get("PREC", "bf16"); a.config = "c2"
dev = torch.device("cuda")
tmpl = shapes.sttran_template() if a.arch == "sttran" else shapes.dsg_template()
tr = Trainer({k: v.to(dev) for k, v in synth.make_state_dict(tmpl, 0).items()}, "sgdet", a.arch, a.precision, device=dev)
host = M.collate(bench.make_videos(a, 0, a.videos), "sgdet", pin=True)
res = M.upload(host, dev, rasterise=False)
def step():
    b = M.Batch(); b.__dict__.update(res.__dict__)
    return tr.step(b)
for _ in range(3): step()
torch.cuda.synchronize()
for _ in range(2):
    t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
    print(f"host time of one step {1e3*(t1-t0):.2f} ms, step wall {1e3*(t2-t0):.2f} ms")
t0 = time.perf_counter(); plan = M.make_plan(res, dev, "sgdet", a.arch == "dsg", with_labels=True); t1 = time.perf_counter()
print(f"make_plan(+labels) {1e3*(t1-t0):.2f} ms")
t0 = time.perf_counter(); la = M.label_arrays(res); t1 = time.perf_counter()
print(f"label_arrays {1e3*(t1-t0):.2f} ms")
torch.cuda.synchronize()
t0 = time.perf_counter(); loss, out = tr.forward_backward(res, plan); t1 = time.perf_counter(); tr.optimizer_step(); t2 = time.perf_counter()
torch.cuda.synchronize()
print(f"forward_backward host {1e3*(t1-t0):.2f} ms, optimizer_step host {1e3*(t2-t1):.2f} ms")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
ev0.record()
for _ in range(5): step()
ev1.record(); torch.cuda.synchronize()
print(f"steady state {ev0.elapsed_time(ev1)/5:.2f} ms/step")
recs = bench.profile_one_step(step)
tot = sum(r[1] for r in recs)
print(f"sum of per-call event times: {tot:.2f} ms over {len(recs)} calls")
agg = {}
for name, ms, flops, units, m, n, k, dt in recs:
    c = agg.setdefault(name, [0, 0.0]); c[0] += 1; c[1] += ms
for k, (n, ms) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f"{ms:8.3f} ms {100*ms/tot:5.1f}%  n={n:4d}  {k}")
print("--- top individual GEMMs")
g = {}
for name, ms, flops, units, m, n, k, dt in recs:
    if name in ("nlv_gemm", "nlv_conv3x3_dgrad"):
        c = g.setdefault((m, n, k, dt), [0, 0.0, flops]); c[0] += 1; c[1] += ms
for (m, n, k, dt), (cnt, ms, fl) in sorted(g.items(), key=lambda kv: -kv[1][1])[:30]:
    print(f"{ms:8.3f} ms n={cnt:3d} m={m} n={n} k={k} dt={dt}  {fl*cnt/ms/1e9:8.1f} TFLOP/s")
