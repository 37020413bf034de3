#!/bin/bash
# A/B runs of bench.py over experiment knobs (cape_set_tuning): usage tools/gpu_ab.sh "<tune1>" "<tune2>" ...  ("-" = defaults)
mkdir -p gpurun_out
for t in "$@"; do
  arg=""; [ "$t" != "-" ] && arg="--tune $t"
  echo "== tune $t"
  timeout 300 python bench.py --config ${CFG:-c3} --steps 20 --warmup 5 --no-cpu-baseline --no-profile $arg 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: b=json.loads(l)
    except Exception: continue
    print('   ms_per_step %.3f  value %.1f  loss %s' % (b['ms_per_step'], b['value'], {k: round(v,4) for k,v in b.get('loss',{}).items()}))
"
done
