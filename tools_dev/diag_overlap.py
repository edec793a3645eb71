import sys, os, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nlvsgg_b200 import ops
dev = torch.device("cuda")
src = torch.empty(1200_000_000, dtype=torch.float32).pin_memory()   # 4.8 GB
dst = torch.empty_like(src, device=dev)
s = torch.cuda.Stream()
def t(fn, n=3):
    fn(); torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(n): fn()
    torch.cuda.synchronize(); return (time.perf_counter() - t0) / n * 1e3
def copy():
    with torch.cuda.stream(s): dst.copy_(src, non_blocking=True)
a = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16); b = torch.randn(8192, 8192, device=dev, dtype=torch.bfloat16)
out = torch.empty(8192, 8192, device=dev)
def mm_torch():
    for _ in range(60): torch.matmul(a, b)
def mm_mine():
    for _ in range(60): ops.gemm(a, b, out)
x = torch.randn(64_000_000, device=dev)
def elem_torch():
    for _ in range(150): x.add_(1.0)
def small_h2d():
    for _ in range(20): torch.zeros(16, dtype=torch.int32).pin_memory().to(dev, non_blocking=True)
print("copy alone %.1f" % t(copy))
for name, fn in (("torch matmul", mm_torch), ("my gemm", mm_mine), ("torch elementwise", elem_torch)):
    alone = t(fn)
    both = t(lambda: (copy(), fn()))
    print(f"{name}: alone {alone:.1f}  with copy {both:.1f}")
