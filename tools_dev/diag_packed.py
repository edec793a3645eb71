"""Dev: where does a resident packed-input step spend its time (host vs device), dropout on / off."""
import cProfile, pstats, os, sys, time, tempfile
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import bench
from nlvsgg_b200 import _C, featfile, model as M, shapes, synth
from nlvsgg_b200.trainer import Trainer

class A: pass
a = A(); a.videos = 64; a.frames = 30; a.boxes = 7; a.arch = "sttran"; a.precision = "bf16"; a.config = "c2"
dev = torch.device("cuda")
sd = synth.make_state_dict(shapes.sttran_template(), 0)
entries = bench.make_videos(a, 0, a.videos, with_gt=True)
tmpdir = tempfile.mkdtemp(prefix="nlv_diag_", dir="/dev/shm")
paths = featfile.write_videos(tmpdir, entries)
host = featfile.Loader(pin=True, depth=1).load(paths)
res = M.upload(host, dev, rasterise=False)
for p_drop in (0.0, 0.1):
    tr = Trainer({k: v.to(dev) for k, v in sd.items()}, "sgdet", "sttran", "bf16", device=dev, dropout=p_drop)
    def step():
        b = M.Batch(); b.__dict__.update(res.__dict__)
        return tr.step(b)
    for _ in range(3): step()
    torch.cuda.synchronize()
    for i in range(6):
        t0 = time.perf_counter(); step(); t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"p={p_drop} step {i}: host {1e3*(t1-t0):.2f} ms, wall {1e3*(t2-t0):.2f} ms", flush=True)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    ev0.record()
    for _ in range(10): step()
    ev1.record(); torch.cuda.synchronize()
    print(f"p={p_drop} steady {ev0.elapsed_time(ev1)/10:.2f} ms/step", flush=True)
    pr = cProfile.Profile(); pr.enable()
    for _ in range(5): step()
    torch.cuda.synchronize()
    pr.disable()
    pstats.Stats(pr).sort_stats("cumulative").print_stats(18)
    del tr
    torch.cuda.empty_cache()
