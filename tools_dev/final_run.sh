set -x
timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -3 > gpurun_out/final_tests.log
timeout 600 python bench.py > gpurun_out/final_bench_n1.log 2>&1
timeout 600 python bench.py --impl reference > gpurun_out/final_bench_ref.log 2>&1
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > gpurun_out/final_smoke.log 2>&1
timeout 300 python tools_dev/bench_eval.py > gpurun_out/final_bench_eval.log 2>&1
bash tools_dev/ncu_capture.sh launches > /dev/null 2>&1
tail -3 gpurun_out/final_tests.log; tail -1 gpurun_out/final_bench_n1.log | cut -c1-300; tail -1 gpurun_out/final_bench_ref.log | cut -c1-300; tail -2 gpurun_out/final_smoke.log; tail -3 gpurun_out/final_bench_eval.log
