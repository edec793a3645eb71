import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from nlvsgg_b200 import model as M, shapes, synth
from nlvsgg_b200.trainer import Trainer
class A: pass
a = A(); a.videos = 64; a.frames = 30; a.boxes = 7; a.arch = "sttran"; a.precision = "bf16"; a.config = "c2"
dev = torch.device("cuda")
tr = Trainer({k: v.to(dev) for k, v in synth.make_state_dict(shapes.sttran_template(), 0).items()}, "sgdet", "sttran", "bf16", device=dev)
host = M.collate(bench.make_videos(a, 0, a.videos), "sgdet", pin=True)
print("pinned:", host.union_feat.is_pinned(), host.features.is_pinned(), "bytes", M.input_bytes(host))
def t(fn, n=3):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
res = M.upload(host, dev)
print("H2D only (main stream): %.1f ms" % t(lambda: M.upload(host, dev, rasterise=False)))
s = torch.cuda.Stream()
def side():
    with torch.cuda.stream(s): M.upload(host, dev, rasterise=False)
print("H2D only (side stream): %.1f ms" % t(side))
def comp():
    b = M.Batch(); b.__dict__.update(res.__dict__); tr.step(M.ensure_masks(b))
for _ in range(2): comp()
print("compute only: %.1f ms" % t(comp))
def both():
    with torch.cuda.stream(s): M.upload(host, dev, rasterise=False)
    comp()
print("H2D(side) + compute concurrently: %.1f ms" % t(both))
def pipe(n=4):
    nxt = tr.prefetch(host)
    for i in range(n):
        l, nxt = tr.step_pipelined(nxt, host if i + 1 < n else None)
        l.item()
torch.cuda.synchronize(); t0 = time.perf_counter(); pipe(4); torch.cuda.synchronize()
print("pipelined per step: %.1f ms" % ((time.perf_counter() - t0) / 4 * 1e3))
