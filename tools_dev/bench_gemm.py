"""Dev: the training step's GEMM shapes in isolation (CUDA events, 10 reps each; operands exceed L2 for the large ones)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nlvsgg_b200 import _C, ops
_C.lib()
dev = "cuda"
bf, f32 = torch.bfloat16, torch.float32
Mr, R49, R196 = 22931, 580895, 2323580
# name, m, n, k, a_major, b_major, out dtype, bias, residual(fp32), relu
SHAPES = [
    ("out_proj fwd (f32 out + bias + f32 residual)", Mr, 1936, 1936, 0, 0, f32, True, True, False),
    ("linear1 fwd (bf16 out + bias + relu)", Mr, 2048, 1936, 0, 0, bf, True, False, True),
    ("linear2 fwd (f32 out + bias + residual)", Mr, 1936, 2048, 0, 0, f32, True, True, False),
    ("qk proj fwd (bf16 out + bias)", Mr, 3872, 1936, 0, 0, bf, True, False, False),
    ("d_o = dy W (bf16 out)", Mr, 1936, 1936, 0, 1, bf, False, False, False),
    ("dx = dqkv W (f32 out + residual)", Mr, 1936, 5808, 0, 1, f32, False, True, False),
    ("dW = dy^T x (f32 out)", 1936, 1936, Mr, 1, 1, f32, False, False, False),
    ("dW lin1 = dh^T x", 2048, 1936, Mr, 1, 1, f32, False, False, False),
    ("union 1x1 conv fwd (bf16 out + bias)", R49, 256, 2048, 0, 0, bf, True, False, False),
    ("union conv dW (split-K)", 256, 2048, R49, 1, 1, f32, False, False, False),
    ("conv3x3 fwd (bf16 out + bias + relu)", R49, 256, 1152, 0, 0, bf, True, False, True),
    ("conv3x3 dcol = dy W (bf16 out)", R49, 1152, 256, 0, 1, bf, False, False, False),
    ("conv3x3 dW (split-K)", 256, 1152, R49, 1, 1, f32, False, False, False),
    ("conv7x7 fwd (bf16 out + bias + relu)", R196, 128, 104, 0, 0, bf, True, False, True),
    ("conv7x7 dW (split-K)", 128, 104, R196, 1, 1, f32, False, False, False),
    ("vr_fc fwd", 11855, 512, 12544, 0, 0, bf, True, False, False),
]
only = os.environ.get("ONLY")
tot = 0.0
for name, m, n, k, am, bm, odt, has_bias, has_res, relu in SHAPES:
    if only and only not in name:
        continue
    a = (torch.randn((m, k) if am == 0 else (k, m), device=dev) * 0.1).to(bf)
    b = (torch.randn((n, k) if bm == 0 else (k, n), device=dev) * 0.1).to(bf)
    out = torch.empty(m, n, device=dev, dtype=odt)
    bias = torch.randn(n, device=dev) if has_bias else None
    res = torch.randn(m, n, device=dev) if has_res else None
    f = lambda: ops.gemm(a, b, out, a_major=am, b_major=bm, bias=bias, residual=res, relu=relu)
    for _ in range(3): f()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): f()
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    nbytes = (m * k + n * k) * 2 + m * n * out.element_size() + (m * n * 4 if has_res else 0)
    print(f"{name:46s} m={m:8d} n={n:5d} k={k:8d}  {ms*1e3:8.1f} us  {2.0*m*n*k/ms/1e9:7.1f} TFLOP/s  {nbytes/ms/1e9:6.2f} TB/s algorithmic", flush=True)
    del a, b, out, res
