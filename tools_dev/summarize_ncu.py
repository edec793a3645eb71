"""Summarise the ncu CSVs that tools_dev/ncu_capture.sh brought back in gpurun_out/ into tracked files under profiles/.

  launches_r1.csv      gpu__time_duration.sum of every launch          -> r1_launches_by_kernel.md, r1_launches.csv.gz
  gemm_dram.csv        DRAM bytes + duration of every tcgen05 GEMM      -> gemm_traffic.json (read by bench.py), r1_gemm_dram.md
  gemm_full_raw.csv    --set full, raw page, a window of GEMM launches  -> r1_gemm_ncu.md
  attn_full_raw.csv    --set full, raw page, attention launches         -> r1_attn_ncu.md
"""
import collections, csv, gzip, json, os, re, shutil, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT = os.path.join(ROOT, "profiles")
GO = os.path.join(ROOT, "gpurun_out")
RND = os.environ.get("ROUND", "r2")     # file-name prefix of this round
STEPS_IN_RUN = 7   # bench.py --steps 1 --warmup 1: warm-up, timed, and the 5 instrumented roofline steps (bench.PROFILE_STEPS)
os.makedirs(OUT, exist_ok=True)
UNIT = {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "ns": 1, "nsecond": 1, "us": 1e3, "usecond": 1e3, "ms": 1e6, "msecond": 1e6,
        "s": 1e9, "second": 1e9}


def short(name):
    k = re.sub(r"\(.*", "", name)
    return re.sub(r"^void ", "", k).replace("nlv::<unnamed>::", "").replace("nlv::(anonymous namespace)::", "")


def long_rows(path):
    """ncu --csv log (one row per launch and metric) -> {launch id: {"kernel":..., metric: value in base units}} in launch order."""
    lines = open(path, errors="replace").read().splitlines()
    start = next(i for i, l in enumerate(lines) if l.startswith('"ID"'))
    out = collections.OrderedDict()
    for r in csv.DictReader(lines[start:]):
        try:
            v = float(r["Metric Value"].replace(",", ""))
        except Exception:
            continue
        d = out.setdefault(int(r["ID"]), {"kernel": r["Kernel Name"], "grid": r.get("Grid Size", "")})
        d[r["Metric Name"]] = v * UNIT.get(r["Metric Unit"], 1)
    return list(out.values())


def wide_rows(path):
    """`ncu -i rep --page raw --csv` (one row per launch, one column per metric, second row = units)."""
    rd = list(csv.reader(open(path, errors="replace")))
    if len(rd) < 3:
        return []
    hdr, units = rd[0], rd[1]
    recs = []
    for r in rd[2:]:
        if len(r) < len(hdr):
            continue
        d = {}
        for h, u, x in zip(hdr, units, r):
            try:
                d[h] = float(x.replace(",", "")) * UNIT.get(u, 1)
            except Exception:
                d[h] = x
        recs.append(d)
    return recs


def launches():
    path = os.path.join(GO, f"launches_{RND}.csv")
    rows = long_rows(path)
    # the last step of the run = everything after the second-to-last AdamW launch (AdamW is the final kernel of a step)
    ad = [i for i, d in enumerate(rows) if "adamw_kernel" in d["kernel"]]
    last = rows[ad[-2] + 1: ad[-1] + 1] if len(ad) >= 2 else rows[-(len(rows) // STEPS_IN_RUN):]
    n = len(last)
    agg = collections.defaultdict(lambda: [0, 0.0])
    for d in last:
        a = agg[short(d["kernel"])]
        a[0] += 1; a[1] += d.get("gpu__time_duration.sum", 0.0)
    tot = sum(v for _, v in agg.values())
    md = [f"# ncu launch list, {RND} — one training step (BASELINE C2: 64 videos, 1976 frames, 11,855 pairs, bf16)", "",
          "`ncu --metrics gpu__time_duration.sum --clock-control none` over `bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras`",
          f"(tools_dev/ncu_r2.sh); last step of the run: {n} launches, {tot/1e6:.2f} ms of kernel time.",
          "Per-launch times under ncu are cold-cache and serialised: compare SHARES with bench.py's live CUDA-event numbers, not absolutes.", "",
          "| kernel | launches | ms | share |", "|---|---:|---:|---:|"]
    for k, (c, v) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        md.append(f"| `{k}` | {c} | {v/1e6:.3f} | {100*v/tot:.1f}% |")
    gem = sum(v for k, (c, v) in agg.items() if "gemm_tc_kernel" in k)
    md += ["", f"tcgen05 GEMM share of the step under ncu: {100*gem/tot:.1f}% (bench.py `roofline.kernel_share_of_step` measures the same share live with CUDA events)."]
    open(os.path.join(OUT, f"{RND}_launches_by_kernel.md"), "w").write("\n".join(md) + "\n")
    with open(path, "rb") as f, gzip.open(os.path.join(OUT, f"{RND}_launches.csv.gz"), "wb") as g:
        shutil.copyfileobj(f, g)
    print("\n".join(md[:16]))


def gemm_dram():
    rows = long_rows(os.path.join(GO, f"gemm_dram_{RND}.csv"))
    n = len(rows) // STEPS_IN_RUN
    last = rows[-n:]
    tot_b = sum(d.get("dram__bytes_read.sum", 0) + d.get("dram__bytes_write.sum", 0) for d in last)
    tot_t = sum(d.get("gpu__time_duration.sum", 0) for d in last)
    tp = "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"
    js = {"launches": n, "dram_bytes_per_launch_avg": tot_b / n, "dram_bytes_total_per_step": tot_b, "kernel_time_ns_total_under_ncu": tot_t,
          "tensor_pipe_active_pct_time_weighted": sum(d.get(tp, 0) * d.get("gpu__time_duration.sum", 0) for d in last) / tot_t,
          "tensor_pipe_active_pct_max": max(d.get(tp, 0) for d in last),
          "source": "ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active... "
                    "--clock-control none -k regex:gemm_tc (every tcgen05 GEMM launch of the last training step; tools_dev/ncu_r2.sh)"}
    json.dump(js, open(os.path.join(OUT, "gemm_traffic.json"), "w"), indent=1)
    md = [f"# DRAM traffic of every tcgen05 GEMM launch of one training step ({RND})", "", "```", json.dumps(js, indent=1), "```", "",
          "tensor pipe % is of the 2.25 PFLOP/s nominal peak (100 % = every cycle a tcgen05.mma slice active).", "",
          "| # | template <BN,A_MN,B_MN,MODE> (MODE 2 = cta_group::2 pair) | grid | time us | DRAM read MB | DRAM write MB | tensor pipe % |", "|---:|---|---|---:|---:|---:|---:|"]
    for i, d in enumerate(last):
        t = re.search(r"gemm_tc_kernel<([^>]*)>", d["kernel"])
        md.append(f"| {i} | {t.group(1) if t else '?'} | {d['grid']} | {d.get('gpu__time_duration.sum', 0)/1e3:.1f} | "
                  f"{d.get('dram__bytes_read.sum', 0)/1e6:.1f} | {d.get('dram__bytes_write.sum', 0)/1e6:.1f} | {d.get(tp, 0):.1f} |")
    open(os.path.join(OUT, f"{RND}_gemm_dram.md"), "w").write("\n".join(md) + "\n")
    print(json.dumps(js, indent=1))


WIDE = [("gpu__time_duration.sum", "time us", 1e-3), ("dram__bytes_read.sum", "DRAM rd MB", 1e-6), ("dram__bytes_write.sum", "DRAM wr MB", 1e-6),
        ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM %", 1), ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor %", 1),
        ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM %", 1), ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue %", 1),
        ("sm__warps_active.avg.pct_of_peak_sustained_active", "warps active %", 1), ("launch__registers_per_thread", "regs", 1),
        ("l1tex__t_sector_hit_rate.pct", "L1 hit %", 1), ("lts__t_sector_hit_rate.pct", "L2 hit %", 1),
        ("smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio", "stall long_sb", 1),
        ("smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio", "stall short_sb", 1),
        ("smsp__average_warps_issue_stalled_wait_per_issue_active.ratio", "stall wait", 1)]


def wide_table(src, dst, title, note):
    recs = wide_rows(os.path.join(GO, src))
    if not recs:
        print("no records in", src); return
    cols = [c for c in WIDE if c[0] in recs[0]]
    md = [f"# {title}", "", note, "", "| # | kernel | grid | " + " | ".join(c[1] for c in cols) + " |", "|---:|---|---|" + "---:|" * len(cols)]
    for i, d in enumerate(recs):
        k = short(str(d.get("Kernel Name", "")))
        vals = " | ".join(f"{d[c[0]] * c[2]:.1f}" if isinstance(d.get(c[0]), float) else "-" for c in cols)
        md.append(f"| {i} | `{k[:70]}` | {d.get('Grid Size', '')} | {vals} |")
    open(os.path.join(OUT, dst), "w").write("\n".join(md) + "\n")
    print("\n".join(md[:12]))


if __name__ == "__main__":
    if os.path.exists(os.path.join(GO, f"launches_{RND}.csv")):
        launches()
    if os.path.exists(os.path.join(GO, f"gemm_dram_{RND}.csv")):
        gemm_dram()
    if os.path.exists(os.path.join(GO, f"gemm_full_raw_{RND}.csv")):
        wide_table(f"gemm_full_raw_{RND}.csv", f"{RND}_gemm_ncu.md", f"ncu --set full, a window of tcgen05 GEMM launches of one training step ({RND})",
                   "`ncu --set full --clock-control none --import-source on -k regex:gemm_tc` over a window of the last step's launches (tools_dev/ncu_r2.sh); "
                   "raw page exported on the box.  gemm_tc_kernel<BN, A_MN, B_MN, MODE>: MODE 2 = CTA pair issuing tcgen05.mma.cta_group::2.")
    if os.path.exists(os.path.join(GO, f"attn_full_raw_{RND}.csv")):
        wide_table(f"attn_full_raw_{RND}.csv", f"{RND}_attn_ncu.md", f"ncu --set full, varlen attention kernels of one training step ({RND})",
                   "`ncu --set full --clock-control none --import-source on -k regex:attn_`: forward kernels, the fused single-tile backward and the "
                   "two-kernel backward (which exits at once for segments the fused kernel took).")
    if os.path.exists(os.path.join(GO, f"fin_full_raw_{RND}.csv")):
        wide_table(f"fin_full_raw_{RND}.csv", f"{RND}_fused_kernels_ncu.md",
                   f"ncu --set full, attention / mask-branch / feature-decode kernels of one training step ({RND}, end of round)",
                   "`ncu --set full --clock-control none --import-source on -k regex:'attn_|mask_conv1|bn_apply_maxpool|pool_bn_bwd|union_unpack12' -c 20` "
                   "(tools_dev/final_ncu_r2.sh; first step of the run, cold caches).")
    if os.path.exists(os.path.join(GO, f"tail_full_raw_{RND}.csv")):
        wide_table(f"tail_full_raw_{RND}.csv", f"{RND}_tail_ncu.md", f"ncu --set full, the memory-bound kernels of one training step ({RND})",
                   "`ncu --set full --clock-control none -k regex:'bn_|colsum|layernorm|union_unpack|maxpool|im2col|col2im|gather|split3|convert'` "
                   "(first step of the run).")
