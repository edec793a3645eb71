#!/bin/bash
# Round-1 ncu captures, run on the GPU box through gpurun (one GPU).  Only CSV text comes back in gpurun_out/
# (the .ncu-rep files stay in /tmp on the box: a 72-launch --set full report is > 64 MiB, the gpurun_out limit).
#   bash tools_dev/ncu_capture.sh [launches|gemm|attn|all]
set -x
what=${1:-all}
B="python bench.py --videos 64 --steps 1 --warmup 1 --no-e2e --no-cpu-baseline"   # 3 training steps: warm-up, timed, instrumented
mkdir -p gpurun_out
if [ "$what" = all ] || [ "$what" = launches ]; then
  # every kernel launch of the run with its duration (B200_PROFILING.md: the launch-list pass)
  timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_r1.csv $B \
    > gpurun_out/ncu_launch_bench.log 2>&1
fi
if [ "$what" = all ] || [ "$what" = gemm ]; then
  # DRAM traffic of EVERY tcgen05 GEMM launch (cheap metrics, all three steps; the summariser keeps the last step)
  timeout 700 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active \
    --clock-control none -k regex:gemm_tc --csv --log-file gpurun_out/gemm_dram.csv $B > gpurun_out/ncu_gemm_dram.log 2>&1
  # full section set for a window of the last step's GEMM launches
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:gemm_tc -s ${GEMM_SKIP:-176} -c ${GEMM_COUNT:-8} -f -o /tmp/gemm_full $B \
    > gpurun_out/ncu_gemm_full.log 2>&1
  ncu -i /tmp/gemm_full.ncu-rep --page raw --csv > gpurun_out/gemm_full_raw.csv 2>/dev/null
fi
if [ "$what" = all ] || [ "$what" = attn ]; then
  # attention kernels of the last step: the spatial encoder's three launches are skipped, first decoder layer fwd + bwd kept
  timeout 500 ncu --set full --clock-control none --import-source on -k regex:attn_ -s ${ATTN_SKIP:-24} -c ${ATTN_COUNT:-12} -f -o /tmp/attn_full $B \
    > gpurun_out/ncu_attn_full.log 2>&1
  ncu -i /tmp/attn_full.ncu-rep --page raw --csv > gpurun_out/attn_full_raw.csv 2>/dev/null
fi
ls -la /tmp/*.ncu-rep; du -sh gpurun_out
