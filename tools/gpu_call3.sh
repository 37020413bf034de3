#!/bin/bash
mkdir -p gpurun_out
echo "== tests"
timeout 400 python -m pytest tests/test_gpu_api.py tests/test_gpu_parity.py -q -m gpu -k "generated or train_step_matches or full_batch or chebyshev_forward or groupnorm_decoder" 2>&1 | grep -v Warning | tail -80 > gpurun_out/pytest_gpu3.log
grep -E "passed|failed" gpurun_out/pytest_gpu3.log | tail -3
grep -E "^E  " gpurun_out/pytest_gpu3.log | cut -c1-800 | head -6
run() { tag=$1; shift; echo "== $tag: $*"; env "$@" timeout 200 python bench.py --steps 20 --warmup 5 --no-cpu-baseline 2>/dev/null | python -c "
import sys,json
for l in sys.stdin:
    try: b=json.loads(l)
    except Exception: continue
    print('   ms_per_step %.3f  value %.1f' % (b['ms_per_step'], b['value']))
"; cp gpurun_out/launch_profile_c3.json gpurun_out/lp_$tag.json; }
run default CAPE_NOOP=1
run fwd_contract CAPE_FWD_MODE=contract
run fwd_basis CAPE_FWD_MODE=basis
run dx_contract CAPE_DX_MODE=contract
run dx_basis CAPE_DX_MODE=basis
run ap4 CAPE_B200_LIB=$PWD/cape_b200/libcape_b200_ap4.so
run ap2 CAPE_B200_LIB=$PWD/cape_b200/libcape_b200_ap2.so
