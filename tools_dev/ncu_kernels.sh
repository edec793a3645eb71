#!/bin/bash
# ncu --set full of a few named kernels of the instrumented bench step: bash tools_dev/ncu_kernels.sh '<kernel regex>' <count> <tag>
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras"
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:"$1" -c ${2:-6} -f -o /tmp/k_$3 $B > gpurun_out/ncu_$3.log 2>&1
ncu -i /tmp/k_$3.ncu-rep --page raw --csv > gpurun_out/${3}_raw.csv 2>/dev/null
ncu -i /tmp/k_$3.ncu-rep --page source --csv --print-source sass > gpurun_out/${3}_source.csv 2>/dev/null
gzip -f gpurun_out/${3}_source.csv
ls -la gpurun_out/${3}_*
