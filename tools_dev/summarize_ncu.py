"""Summarise ncu outputs brought back in gpurun_out/ into tracked files under profiles/.
  launches csv (gpu__time_duration.sum per launch)  -> profiles/r1_launches_by_kernel.md (+ copy of the csv, gzipped)
  prof_gemm_step.ncu-rep (--set full, all tcgen05 GEMM launches of one step) -> profiles/gemm_traffic.json + r1_gemm_ncu.md
"""
import collections, csv, gzip, io, json, os, re, shutil, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
os.makedirs(OUT, exist_ok=True)


def launches(path, steps_in_run=2):
    lines = open(path).read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    rows = []
    for r in csv.DictReader(lines[start:]):
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u = r["Metric Unit"]
        v *= {"ns": 1, "us": 1e3, "ms": 1e6, "s": 1e9}.get(u, 1)
        rows.append((int(r["ID"]), r["Kernel Name"], v))
    n = len(rows) // steps_in_run
    last = rows[-n:]          # the last step of the run (timed step or roofline step): warm caches for code, cold for data
    agg = collections.defaultdict(lambda: [0, 0.0])
    for _, name, v in last:
        k = re.sub(r"\(.*", "", name)
        k = re.sub(r"^void ", "", k).replace("nlv::<unnamed>::", "")
        agg[k][0] += 1; agg[k][1] += v
    tot = sum(v for _, v in agg.values())
    md = ["# ncu launch list, round 1 — one training step (BASELINE C2: 64 videos, 1976 frames, 11,855 pairs, bf16)", "",
          "`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --videos 64 --steps 1 --warmup 1`;",
          f"last step of the run: {n} launches, {tot/1e6:.2f} ms of kernel time (serialised, cold-cache: compare SHARES).", "",
          "| kernel | launches | ms | share |", "|---|---:|---:|---:|"]
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        md.append(f"| `{k}` | {c} | {v/1e6:.3f} | {100*v/tot:.1f}% |")
    gem = sum(v for k, (c, v) in agg.items() if "gemm_tc_kernel" in k)
    md += ["", f"tcgen05 GEMM share of the step: {100*gem/tot:.1f}% (bench.py `roofline.kernel_share_of_step` measures the same share live with CUDA events)."]
    open(os.path.join(OUT, "r1_launches_by_kernel.md"), "w").write("\n".join(md) + "\n")
    with open(path, "rb") as f, gzip.open(os.path.join(OUT, "r1_launches.csv.gz"), "wb") as g:
        shutil.copyfileobj(f, g)
    print("\n".join(md[:14]))


def gemm_rep(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rd = list(csv.reader(io.StringIO(out)))
    hdr = rd[0]
    col = {h: i for i, h in enumerate(hdr)}
    want = ["dram__bytes_read.sum", "dram__bytes_write.sum", "gpu__time_duration.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
            "sm__inst_executed_pipe_tensor.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "launch__registers_per_thread",
            "sm__warps_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
    units = rd[1]
    rows = rd[2:]
    def val(r, name):
        i = col.get(name)
        if i is None:
            return None
        try:
            return float(r[i].replace(",", ""))
        except Exception:
            return None
    def scale(name, v):
        u = units[col[name]] if name in col else ""
        m = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "us": 1e3, "usecond": 1e3, "msecond": 1e6, "ms": 1e6, "nsecond": 1, "second": 1e9}
        return v * m.get(u, 1) if v is not None else None
    recs = []
    for r in rows:
        if len(r) < len(hdr):
            continue
        d = {n: scale(n, val(r, n)) for n in want}
        d["kernel"] = r[col["Kernel Name"]] if "Kernel Name" in col else ""
        recs.append(d)
    if not recs:
        print("no records in", path); return
    tot_b = sum((d["dram__bytes_read.sum"] or 0) + (d["dram__bytes_write.sum"] or 0) for d in recs)
    tot_t = sum(d["gpu__time_duration.sum"] or 0 for d in recs)
    tp = [d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] for d in recs if d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] is not None]
    js = {"launches": len(recs), "dram_bytes_per_launch_avg": tot_b / len(recs), "dram_bytes_total": tot_b, "kernel_time_ns_total_under_ncu": tot_t,
          "tensor_pipe_active_pct_time_weighted": (sum((d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"] or 0) * (d["gpu__time_duration.sum"] or 0) for d in recs) / tot_t) if tot_t else None,
          "tensor_pipe_active_pct_max": max(tp) if tp else None,
          "source": "ncu --set full --clock-control none -k regex:gemm_tc_kernel -s 72 -c 72 (all tcgen05 GEMM launches of one training step)"}
    json.dump(js, open(os.path.join(OUT, "gemm_traffic.json"), "w"), indent=1)
    md = ["# ncu --set full, tcgen05 GEMM launches of one training step (round 1)", "", "```", json.dumps(js, indent=1), "```", "",
          "| # | kernel | time us | DRAM MB (r+w) | tensor pipe % | DRAM % | regs |", "|---:|---|---:|---:|---:|---:|---:|"]
    for i, d in enumerate(recs):
        k = re.sub(r"\(.*", "", d["kernel"]).replace("void nlv::<unnamed>::", "")
        md.append(f"| {i} | `{k}` | {(d['gpu__time_duration.sum'] or 0)/1e3:.1f} | {((d['dram__bytes_read.sum'] or 0)+(d['dram__bytes_write.sum'] or 0))/1e6:.1f} | "
                  f"{d['sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active'] or 0:.1f} | {d['gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed'] or 0:.1f} | {int(d['launch__registers_per_thread'] or 0)} |")
    open(os.path.join(OUT, "r1_gemm_ncu.md"), "w").write("\n".join(md) + "\n")
    print(json.dumps(js, indent=1))


if __name__ == "__main__":
    go = os.path.join(ROOT, "gpurun_out")
    if os.path.exists(os.path.join(go, "launches_r1.csv")):
        launches(os.path.join(go, "launches_r1.csv"))
    if os.path.exists(os.path.join(go, "prof_gemm_step.ncu-rep")):
        gemm_rep(os.path.join(go, "prof_gemm_step.ncu-rep"))
