# Round-2 collection run on one GPU (under gpurun): tests, smoke, every bench configuration.  ncu captures: tools_dev/ncu_r2.sh
set -x
timeout 1200 python -m pytest tests -m gpu -q > gpurun_out/r2_final_tests.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2_final_smoke.log 2>&1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
timeout 600 python bench.py --impl reference --steps 1 --warmup 1 > gpurun_out/r2_bench_reference.json 2> gpurun_out/r2_bench_reference.err
for c in c1 c3 c4 c5; do
  timeout 600 python bench.py --config $c --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2_bench_$c.json 2> gpurun_out/r2_bench_$c.err
done
set +x
tail -3 gpurun_out/r2_final_tests.log | cut -c1-300; tail -2 gpurun_out/r2_final_smoke.log; for f in n1 reference c1 c3 c4 c5; do tail -c 300 gpurun_out/r2_bench_$f.json; echo; done
