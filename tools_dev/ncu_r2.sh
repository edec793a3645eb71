#!/bin/bash
# Round-2 ncu captures (one GPU, under gpurun).  Text only comes back (reports stay in /tmp on the box).
#   bash tools_dev/ncu_r2.sh [launches|tail|gemm|attn]
what=${1:-launches}
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras"   # 3 training steps: warm-up, timed, instrumented
mkdir -p gpurun_out
if [ "$what" = launches ]; then
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r2.csv $B > gpurun_out/ncu_launch_r2.log 2>&1
fi
if [ "$what" = tail ]; then
  # the non-GEMM kernels of the first step, full sections
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:'bn_|colsum|layernorm|attn_|union_unpack|maxpool|im2col|col2im|gather|split3|convert' \
    -c ${COUNT:-110} -f -o /tmp/tail_full $B > gpurun_out/ncu_tail_r2.log 2>&1
  ncu -i /tmp/tail_full.ncu-rep --page raw --csv > gpurun_out/tail_full_raw.csv 2>/dev/null
fi
if [ "$what" = gemm ]; then
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s ${GEMM_SKIP:-85} -c ${GEMM_COUNT:-85} -f -o /tmp/gemm_full $B \
    > gpurun_out/ncu_gemm_r2.log 2>&1
  ncu -i /tmp/gemm_full.ncu-rep --page raw --csv > gpurun_out/gemm_full_raw_r2.csv 2>/dev/null
fi
ls -la /tmp/*.ncu-rep 2>/dev/null; du -sh gpurun_out
