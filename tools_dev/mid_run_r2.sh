# mid-round validation: all GPU tests, smoke, every bench configuration (no ncu)
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/r2m_tests.log 2>&1; tail -3 gpurun_out/r2m_tests.log | cut -c1-200
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2m_smoke.log 2>&1; tail -1 gpurun_out/r2m_smoke.log
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2m_bench_n1.json 2> gpurun_out/r2m_bench_n1.err
for c in c1 c3 c4 c5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2m_bench_$c.json 2> gpurun_out/r2m_bench_$c.err
done
python - <<P
import json
for f in ["n1","c1","c3","c4","c5"]:
    try:
        d=json.loads(open("gpurun_out/r2m_bench_%s.json"%f).read().strip().splitlines()[-1])
        e=d.get("e2e") or {}
        print(f, d["metric"], round(d["value"],1), round(d["ms_per_step"],3), "e2e", e.get("value"), e.get("ms_per_step"), e.get("h2d_alone_gb_per_s_per_gpu"), "frac", (d.get("roofline") or {}).get("frac"), d.get("clocks",{}).get("sm_mhz"))
        if f=="n1":
            print({k:(round(v["frac"],3), round(v["ms_per_step"],3)) for k,v in d["roofline_extra"].items() if isinstance(v,dict)})
            print(d.get("cpu_baseline"), d.get("parity"))
    except Exception as ex: print(f,"ERR",ex)
P
