set -x
B="python bench.py --videos 64 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv $B > gpurun_out/ncu_launch_bench.log 2>&1
timeout 500 ncu --set full --clock-control none --import-source on -k regex:attn -s 12 -c 12 -f -o gpurun_out/attn_r1 $B > gpurun_out/ncu_attn.log 2>&1
ncu -i gpurun_out/attn_r1.ncu-rep --page raw --csv > gpurun_out/attn_raw.csv 2>/dev/null
ls -la gpurun_out/attn_r1.ncu-rep
[ $(stat -c %s gpurun_out/attn_r1.ncu-rep) -gt 30000000 ] && rm gpurun_out/attn_r1.ncu-rep
timeout 600 ncu --set full --clock-control none -k regex:gemm_tc -s 72 -c 72 -f -o /tmp/gemm_r1 $B > gpurun_out/ncu_gemm.log 2>&1
ncu -i /tmp/gemm_r1.ncu-rep --page raw --csv > gpurun_out/gemm_raw.csv 2>/dev/null
ls -la /tmp/gemm_r1.ncu-rep; du -sh gpurun_out
