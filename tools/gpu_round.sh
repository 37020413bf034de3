#!/bin/bash
# One GPU-box visit: parity checks, tests, bench (+ optional ncu).  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L | head -1
echo "== gpu_check"; timeout 900 python tests/gpu_check.py golden tc cheb step step_n5 step_gn 2>&1 | grep -v "Warn\|warn\|return torch\|out = {" > gpurun_out/gpu_check.log; grep -c FAIL gpurun_out/gpu_check.log; grep -v "param-update\|  grad " gpurun_out/gpu_check.log | tail -60; grep FAIL gpurun_out/gpu_check.log | head -30
echo "== pytest -m gpu"; timeout 1200 python -m pytest tests -q -m gpu 2>&1 > gpurun_out/pytest_gpu.log; tail -5 gpurun_out/pytest_gpu.log
echo "== bench"; timeout 900 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench.err | tee gpurun_out/bench.json | cut -c1-400; tail -3 gpurun_out/bench.err
if [ -n "$AB_ENV$AB_ARGS" ]; then
echo "== bench A/B: $AB_ENV $AB_ARGS"; env $AB_ENV timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile $AB_ARGS 2> gpurun_out/bench_ab.err | tee gpurun_out/bench_ab.json | cut -c1-400
fi
if [ "$1" == "ncu" ]; then
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 450 -c 900 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-profile > gpurun_out/ncu_bench.log 2>&1
tail -2 gpurun_out/ncu_bench.log | cut -c1-200; wc -l gpurun_out/launches.csv
echo "== ncu full (wide conv, narrow conv, dense dW)"
for spec in "conv_wide:ellconv_tc_kernel:12" "conv_narrow:ellconv_tc2_kernel:20" "dw_dense:dw_dense_kernel:16"; do
  name="${spec%%:*}"; rest="${spec#*:}"; kern="${rest%%:*}"; skip="${rest#*:}"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c 4 -f -o gpurun_out/prof_$name \
     python bench.py --steps 1 --warmup 0 --no-graph --no-cpu-baseline --no-profile > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log | cut -c1-150
done
fi
ls -la gpurun_out | head -30
