#!/bin/bash
# Round-2 ncu captures (one GPU, under gpurun).  Text only comes back (reports stay in /tmp on the box).
#   bash tools_dev/ncu_r2.sh [launches|gemm|attn|tail|all]     then here: ROUND=r2 python tools_dev/summarize_ncu.py
what=${1:-all}
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras"   # 3 training steps: warm-up, timed, instrumented
mkdir -p gpurun_out
if [ "$what" = launches ] || [ "$what" = all ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/ncu_launch_r2.log 2>&1
fi
if [ "$what" = gemm ] || [ "$what" = all ]; then
  # DRAM traffic + tensor-pipe activity of EVERY tcgen05 GEMM launch (cheap metrics; the summariser keeps the last step)
  timeout 700 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:gemm_tc --csv --log-file gpurun_out/gemm_dram_r2.csv $B > gpurun_out/ncu_gemm_dram_r2.log 2>&1
  # full sections for a window of the last step's launches (forward transformer GEMMs + a few backward ones)
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s ${GEMM_SKIP:-190} -c ${GEMM_COUNT:-12} -f -o /tmp/gemm_full $B \
    > gpurun_out/ncu_gemm_r2.log 2>&1
  ncu -i /tmp/gemm_full.ncu-rep --page raw --csv > gpurun_out/gemm_full_raw_r2.csv 2>/dev/null
fi
if [ "$what" = attn ] || [ "$what" = all ]; then
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:attn_ -c ${ATTN_COUNT:-16} -f -o /tmp/attn_full $B > gpurun_out/ncu_attn_r2.log 2>&1
  ncu -i /tmp/attn_full.ncu-rep --page raw --csv > gpurun_out/attn_full_raw_r2.csv 2>/dev/null
fi
if [ "$what" = tail ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bn_|colsum|layernorm|union_unpack|maxpool|im2col|col2im|gather|split3|convert' \
    -c ${COUNT:-110} -f -o /tmp/tail_full $B > gpurun_out/ncu_tail_r2.log 2>&1
  ncu -i /tmp/tail_full.ncu-rep --page raw --csv > gpurun_out/tail_full_raw_r2.csv 2>/dev/null
fi
ls -la /tmp/*.ncu-rep 2>/dev/null; du -sh gpurun_out
