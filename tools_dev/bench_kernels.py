"""Dev: stand-alone timings of the memory-bound kernels at the C2 shapes (CUDA events, 20 reps, operands larger than L2)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from nlvsgg_b200 import _C, ops
_C.lib()
dev = "cuda"

def timeit(fn, n=20):
    for _ in range(3): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / n

def report(name, ms, nbytes):
    print(f"{name:46s} {ms*1e3:8.1f} us  {nbytes/ms/1e9:7.2f} TB/s ({nbytes/ms/1e9/6.553*100:5.1f}% of 6553 GB/s)", flush=True)

Mr, D = 22931, 1936
x = torch.randn(Mr, D, device=dev); dy = torch.randn(Mr, D, device=dev)
w = torch.randn(D, device=dev); b = torch.randn(D, device=dev)
_, _, mean, rstd = ops.layernorm_fwd(x, w, b)
d = _C.Dropout.make(0.1, 5, 2)
report("layernorm_bwd (dx + param kernels), bf16 copy", timeit(lambda: ops.layernorm_bwd(dy, x, mean, rstd, w, dx2_dtype=torch.bfloat16, drop=d)), Mr * D * 14)
report("layernorm_bwd_fused (+ colsum), bf16 copy", timeit(lambda: ops.layernorm_bwd_fused(dy, x, mean, rstd, w, dx2_dtype=torch.bfloat16, drop=d)), Mr * D * 14)
report("layernorm_fwd, bf16 copy", timeit(lambda: ops.layernorm_fwd(x, w, b, y2_dtype=torch.bfloat16)), Mr * D * 10)
report("colsum fp32 [22931,1936]", timeit(lambda: ops.colsum(dy)), Mr * D * 4)
h = torch.randn(Mr, 2048, device=dev).bfloat16()
report("colsum bf16 [22931,2048]", timeit(lambda: ops.colsum(h)), Mr * 2048 * 2)
q = torch.randn(Mr, 5808, device=dev).bfloat16()
report("colsum bf16 [22931,5808]", timeit(lambda: ops.colsum(q)), Mr * 5808 * 2)
# attention: temporal-decoder shape (2-frame windows of ~12 rows), bf16, with dropout masks
import numpy as np
from nlvsgg_b200.plan import work_items
rng = np.random.default_rng(0)
lens = []
while sum(lens) < Mr:
    lens.append(int(rng.integers(8, 17)))
lens[-1] -= sum(lens) - Mr
if lens[-1] <= 0: lens.pop(); lens[-1] += Mr - sum(lens)
starts = np.concatenate(([0], np.cumsum(lens)))[:-1]
work = torch.from_numpy(work_items(starts, np.asarray(lens))).cuda()
qkv = (torch.randn(Mr, 3 * D, device=dev) * 0.5).bfloat16()
qq, kk, vv = qkv[:, :D], qkv[:, D:2 * D], qkv[:, 2 * D:]
dd = _C.Dropout.make(0.1, 9, 3)
o, lse = ops.attn_fwd(qq, kk, vv, 242, 8, work, work.shape[0], torch.bfloat16, drop=dd)
report("attention fwd (bf16, %d segments)" % len(lens), timeit(lambda: ops.attn_fwd(qq, kk, vv, 242, 8, work, work.shape[0], torch.bfloat16, drop=dd)), Mr * D * 2 * 4)
do = torch.randn(Mr, D, device=dev).bfloat16()
dqkv = torch.empty_like(qkv)
report("attention bwd (bf16)", timeit(lambda: ops.attn_bwd(qq, kk, vv, o, do, lse, 242, 8, work, work.shape[0], dqkv[:, :D], dqkv[:, D:2 * D], dqkv[:, 2 * D:], drop=dd)), Mr * D * 2 * 8)
if os.environ.get("ATTN_ONLY"): sys.exit(0)
# conv stack BatchNorm shapes: 64 videos, R = 11855 pairs
R, nv = 11855, 64
for rows_per_pair, C in ((196, 128), (49, 256)):
    rows = R * rows_per_pair
    xx = torch.randn(rows, C, device=dev).bfloat16()
    per = torch.full((nv,), R // nv, dtype=torch.int64); per[: R % nv] += 1
    seg = torch.cat((torch.zeros(1, dtype=torch.int64), torch.cumsum(per * rows_per_pair, 0))).to(torch.int32).to(dev)
    row_seg = torch.repeat_interleave(torch.arange(nv, dtype=torch.int32), per).to(dev)
    rm, rv = torch.zeros(C, device=dev), torch.ones(C, device=dev)
    wc, bc = torch.randn(C, device=dev), torch.randn(C, device=dev)
    mean_c, var_c = ops.bn_stats(xx, seg, nv, C, 0.01, rm, rv)
    report(f"bn_stats bf16 [{rows},{C}]", timeit(lambda: ops.bn_stats(xx, seg, nv, C, 0.01, rm, rv)), rows * C * 2)
    y, _ = ops.bn_apply(xx, row_seg, mean_c, var_c, wc, bc, False, out_dtype=torch.bfloat16, row_div=rows_per_pair)
    report(f"bn_apply bf16 [{rows},{C}]", timeit(lambda: ops.bn_apply(xx, row_seg, mean_c, var_c, wc, bc, False, out=y, row_div=rows_per_pair)), rows * C * 4)
    gy = torch.randn(rows, C, device=dev).bfloat16()
    report(f"bn_bwd bf16 (sums + apply) [{rows},{C}]", timeit(lambda: ops.bn_bwd(gy, xx, None, seg, row_seg, nv, mean_c, var_c, wc, True, dx_dtype=torch.bfloat16, gate_by_x=True, row_div=rows_per_pair)), rows * C * 2 * 5)
    del xx, y, gy
