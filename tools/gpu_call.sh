#!/bin/bash
# One GPU-box visit under a tight budget: the GPU test suite, the three bench configs, the ncu launch list.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
nvidia-smi -L | head -1
echo "== pytest -m gpu"; timeout 1300 python -m pytest tests -q -m gpu --timeout 700 --durations=12 2>&1 | tail -60 > gpurun_out/pytest_gpu.log; tail -6 gpurun_out/pytest_gpu.log
echo "== bench c3 (default)"; timeout 400 python bench.py --config c3 --steps 20 --warmup 5 2> gpurun_out/bench_c3.err > gpurun_out/bench_c3.json; cut -c1-330 gpurun_out/bench_c3.json
for c in c2 c5; do
  echo "== bench $c"; timeout 400 python bench.py --config $c --steps 20 --warmup 5 2> gpurun_out/bench_$c.err > gpurun_out/bench_$c.json; cut -c1-260 gpurun_out/bench_$c.json
done
cp gpurun_out/bench_c3.json gpurun_out/bench.json; cp gpurun_out/launch_profile_c3.json gpurun_out/launch_profile.json
echo "== ncu launch list"
timeout 500 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 1100 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-profile > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200; wc -l gpurun_out/launches.csv
ls gpurun_out | head -30
