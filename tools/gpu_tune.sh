#!/bin/bash
mkdir -p gpurun_out
for t in "1=0" "1=1"; do
  echo "== tune $t"; timeout 300 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --tune "$t" 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read()); print(d['value'], d['ms_per_step'], {k:round(v['ms'],2) for k,v in d['roofline']['families'].items()})"
done
