#!/bin/bash
# ncu --set full of the GEMM launches matching a demangled-name regex (e.g. the conv3x3 data-gradient product <256, 0, 1, 1>)
B="python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu-baseline --no-extras"
timeout 600 ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$1" -c ${2:-2} -f -o /tmp/k_$3 $B > gpurun_out/ncu_$3.log 2>&1
ncu -i /tmp/k_$3.ncu-rep --page raw --csv > gpurun_out/${3}_raw.csv 2>/dev/null
ls -la gpurun_out/${3}_*; tail -3 gpurun_out/ncu_$3.log
