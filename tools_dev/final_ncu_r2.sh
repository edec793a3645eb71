# final ncu evidence of round 2 (one GPU): launch list, GEMM DRAM / tensor metrics, --set full of attention and mask-branch kernels
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras"
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/ncu_launch_r2.log 2>&1
timeout 300 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
  --clock-control none -k regex:gemm_tc --csv --log-file gpurun_out/gemm_dram_r2.csv $B > gpurun_out/ncu_gemm_dram_r2.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:'attn_|mask_conv1|bn_apply_maxpool|pool_bn_bwd|union_unpack12' -c 20 -f -o /tmp/k_fin $B > gpurun_out/ncu_fin.log 2>&1
ncu -i /tmp/k_fin.ncu-rep --page raw --csv > gpurun_out/fin_full_raw_r2.csv 2>/dev/null
timeout 400 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_n1.json 2> gpurun_out/r2_bench_n1.err
ls -la gpurun_out/launches_r2.csv gpurun_out/gemm_dram_r2.csv gpurun_out/fin_full_raw_r2.csv; tail -c 600 gpurun_out/r2_bench_n1.json
