"""Dev: packed-input end-to-end step: copy alone, compute alone, both, pipelined."""
import os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, bench
from nlvsgg_b200 import featfile, model as M, shapes, synth
from nlvsgg_b200.trainer import Trainer
class A: pass
a = A(); a.videos = 64; a.frames = 30; a.boxes = 7; a.arch = "sttran"; a.precision = "bf16"; a.config = "c2"
dev = torch.device("cuda")
tr = Trainer({k: v.to(dev) for k, v in synth.make_state_dict(shapes.sttran_template(), 0).items()}, "sgdet", "sttran", "bf16", device=dev, dropout=0.1)
entries = bench.make_videos(a, 0, a.videos, with_gt=True)
tmpdir = tempfile.mkdtemp(prefix="nlv_diag_", dir="/dev/shm")
host = featfile.Loader(pin=True, depth=1).load(featfile.write_videos(tmpdir, entries))
for k in M.TENSOR_KEYS:
    t = getattr(host, k, None)
    if t is not None: print(k, tuple(t.shape), t.dtype, "pinned" if t.is_pinned() else "PAGEABLE", t.numel() * t.element_size())
print("bytes", M.input_bytes(host))
def t(fn, n=5):
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
res = M.upload(host, dev, rasterise=False)
print("H2D only (main stream): %.1f ms" % t(lambda: M.upload(host, dev, rasterise=False)))
s = torch.cuda.Stream()
def side():
    with torch.cuda.stream(s): M.upload(host, dev, rasterise=False)
print("H2D only (side stream): %.1f ms" % t(side))
def comp():
    b = M.Batch(); b.__dict__.update(res.__dict__); tr.step(b)
for _ in range(3): comp()
print("compute only: %.1f ms" % t(comp))
def both():
    with torch.cuda.stream(s): M.upload(host, dev, rasterise=False)
    comp()
print("H2D(side) + compute concurrently: %.1f ms" % t(both))
def pipe(n):
    nxt = tr.prefetch(host)
    for i in range(n):
        t0 = time.perf_counter()
        l, nxt = tr.step_pipelined(nxt, host if i + 1 < n else None)
        t1 = time.perf_counter()
        l.item()
        t2 = time.perf_counter()
        print("   step %d: host enqueue %.2f ms, wait for loss %.2f ms" % (i, 1e3 * (t1 - t0), 1e3 * (t2 - t1)))
pipe(3)
torch.cuda.synchronize(); t0 = time.perf_counter(); pipe(8); torch.cuda.synchronize()
print("pipelined per step: %.1f ms" % ((time.perf_counter() - t0) / 8 * 1e3))
def pipe2(n):
    """compute of step i enqueued BEFORE the host prepares / enqueues the copies of step i+1"""
    nxt = tr.prefetch(host)
    for i in range(n):
        b, plan, ev = nxt
        torch.cuda.current_stream().wait_event(ev)
        tr.forward_backward(b, plan)
        tr.optimizer_step()
        tk = tr.last_ticket
        nxt = tr.prefetch(host) if i + 1 < n else None
        tr.loss_value(tk)
pipe2(3)
torch.cuda.synchronize(); t0 = time.perf_counter(); pipe2(8); torch.cuda.synchronize()
print("compute-first pipelined per step: %.1f ms" % ((time.perf_counter() - t0) / 8 * 1e3))
def pipe3(n):
    """lagged loss read, current ordering"""
    nxt = tr.prefetch(host); tk = None
    for i in range(n):
        l, nxt = tr.step_pipelined(nxt, host if i + 1 < n else None)
        if tk is not None: tr.loss_value(tk)
        tk = tr.last_ticket
    tr.loss_value(tk)
pipe3(3)
torch.cuda.synchronize(); t0 = time.perf_counter(); pipe3(8); torch.cuda.synchronize()
print("lagged-loss pipelined per step: %.1f ms" % ((time.perf_counter() - t0) / 8 * 1e3))
def pipe4(n):
    """no copies at all: the loop overhead alone"""
    tk = None
    for i in range(n):
        bb = M.Batch(); bb.__dict__.update(res.__dict__)
        tr.step(bb)
        if tk is not None: tr.loss_value(tk)
        tk = tr.last_ticket
    tr.loss_value(tk)
pipe4(3)
torch.cuda.synchronize(); t0 = time.perf_counter(); pipe4(8); torch.cuda.synchronize()
print("resident loop with lagged loss per step: %.1f ms" % ((time.perf_counter() - t0) / 8 * 1e3))
print("mem allocated %.1f GB reserved %.1f GB" % (torch.cuda.memory_allocated() / 1e9, torch.cuda.memory_reserved() / 1e9))
# ---- interference test: the resident loop with an unrelated 1.2 GB H2D copy running on a side stream every step
big_h = host.union_feat
big_d = torch.empty_like(big_h, device=dev)
cs = torch.cuda.Stream()
def pipe5(n, copy=True):
    tk = None
    for i in range(n):
        if copy:
            with torch.cuda.stream(cs):
                big_d.copy_(big_h, non_blocking=True)
        bb = M.Batch(); bb.__dict__.update(res.__dict__)
        tr.step(bb)
        if tk is not None: tr.loss_value(tk)
        tk = tr.last_ticket
    tr.loss_value(tk)
for cp in (False, True, False, True):
    pipe5(3, cp)
    torch.cuda.synchronize(); t0 = time.perf_counter(); pipe5(10, cp); torch.cuda.synchronize()
    print("resident loop, unrelated H2D on a side stream = %s: %.1f ms/step" % (cp, (time.perf_counter() - t0) / 10 * 1e3))
# copy alone into a preallocated buffer
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(5):
    with torch.cuda.stream(cs): big_d.copy_(big_h, non_blocking=True)
torch.cuda.synchronize()
print("1.19 GB H2D alone: %.1f ms (%.1f GB/s)" % ((time.perf_counter() - t0) / 5 * 1e3, big_h.numel() * 2 / ((time.perf_counter() - t0) / 5) / 1e9))
