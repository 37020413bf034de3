#!/bin/bash
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:ellconv_dw_tc_kernel -s 40 -c 6 -o gpurun_out/prof_dw_tc \
   python bench.py --steps 1 --warmup 0 --no-graph --no-cpu-baseline --no-profile > gpurun_out/ncu_dw.log 2>&1
tail -2 gpurun_out/ncu_dw.log
