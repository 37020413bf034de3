#!/bin/bash
# Short visit: the two tests that failed in the first one (full output), then A/B of the launch-shape knobs.
mkdir -p gpurun_out
echo "== failing tests"
timeout 400 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py -q -m gpu -k "generated or adam" 2>&1 | grep -v Warning | tail -120 > gpurun_out/pytest_gpu2.log
grep -E "passed|failed" gpurun_out/pytest_gpu2.log | tail -3
grep -E "^E  " gpurun_out/pytest_gpu2.log | cut -c1-1500 | head -20
ab() { echo "== A/B $*"; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline --no-profile 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: b=json.loads(l)
    except Exception: continue
    print('   ms_per_step %.3f  value %.1f  e2e %.1f' % (b['ms_per_step'], b['value'], b['e2e']['value']))
"; }
ab CAPE_WPREP_BLOCKS=16
ab CAPE_WPREP_BLOCKS=128
ab CAPE_WPREP_BLOCKS=128 CAPE_SMALL_BLOCKS=324
ab CAPE_WPREP_BLOCKS=128 CAPE_ASYNC_DW=0
