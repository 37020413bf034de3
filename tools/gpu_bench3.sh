#!/bin/bash
# The three bench configs (gpurun_out/bench_<c>.json + launch_profile_<c>.json), nothing else.
mkdir -p gpurun_out
for c in c3 c2 c5; do
  echo "== bench $c"; timeout 150 python bench.py --config $c --steps 20 --warmup 5 2> gpurun_out/bench_$c.err > gpurun_out/bench_$c.json; cut -c1-200 gpurun_out/bench_$c.json
done
cp gpurun_out/bench_c3.json gpurun_out/bench.json; cp gpurun_out/launch_profile_c3.json gpurun_out/launch_profile.json
