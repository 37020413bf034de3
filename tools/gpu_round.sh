#!/bin/bash
# One GPU-box visit for the round's evidence: tests, the three bench configs, ncu launch list and full captures of the
# top kernels.  Everything lands in gpurun_out/; tools/summarize_profiles.py condenses it into profiles/<tag>_*.
mkdir -p gpurun_out
nvidia-smi -L | head -1
if [ "$1" != "nopytest" ]; then
echo "== pytest -m gpu"; timeout 2400 python -m pytest tests -q -m gpu 2>&1 | tail -8 > gpurun_out/pytest_gpu.log; tail -4 gpurun_out/pytest_gpu.log
fi
echo "== smoke"; timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
for c in c3 c2 c5; do
  echo "== bench $c"; timeout 900 python bench.py --config $c --steps 20 --warmup 5 2> gpurun_out/bench_$c.err > gpurun_out/bench_$c.json; cut -c1-260 gpurun_out/bench_$c.json
done
cp gpurun_out/bench_c3.json gpurun_out/bench.json; cp gpurun_out/launch_profile_c3.json gpurun_out/launch_profile.json
echo "== reference arm (16 meshes/step sample for the record; the driver runs the full one)"
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 --batch 16 2>/dev/null | cut -c1-300 | tee gpurun_out/bench_reference_sample.json
echo "== ncu launch list"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -s 500 -c 1100 --csv --log-file gpurun_out/launches.csv \
   python bench.py --steps 1 --warmup 1 --no-graph --no-cpu-baseline --no-profile > gpurun_out/ncu_bench.log 2>&1
tail -1 gpurun_out/ncu_bench.log | cut -c1-200; wc -l gpurun_out/launches.csv
echo "== ncu full captures"
rm -f gpurun_out/*.ncu-rep
for spec in "gemm_tc:gemm_tc_kernel:6:6" "apply:apply_kernel:4:4" "conv_wide:ellconv_tc_kernel:6:4" "conv_narrow:ellconv_tc2_kernel:10:4" "dw_dense:dw_dense_kernel:16:4"; do
  IFS=: read name kern skip cnt <<< "$spec"
  timeout 600 ncu --set full --clock-control none --import-source on -k regex:$kern -s $skip -c $cnt -f -o gpurun_out/prof_$name \
     python bench.py --steps 1 --warmup 0 --no-graph --no-cpu-baseline --no-profile > gpurun_out/ncu_$name.log 2>&1
  tail -1 gpurun_out/ncu_$name.log | cut -c1-120
done
ls -la gpurun_out | head -40
