#!/bin/bash
# A/B benches on one box: each argument is "ENV=.. ENV=..|bench args"; prints meshes/s for each
mkdir -p gpurun_out
for spec in "$@"; do
  envs="${spec%%|*}"; args="${spec#*|}"
  r=$(env $envs timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-profile $args 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('%.1f meshes/s  %.3f ms' % (d['value'], d['ms_per_step']))" 2>&1 | tail -1)
  echo "[$envs|$args] $r"
done
