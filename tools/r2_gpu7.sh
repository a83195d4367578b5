#!/bin/bash
# round-2 run 7: re-validate the failed / new GPU tests, then per-module ncu launch lists
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.build()" > gpurun_out/r7_build.log 2>&1
timeout 1500 python -m pytest tests/test_gpu_profiles_ext.py tests/test_gpu_profile.py tests/test_gpu_field.py tests/test_gpu_fpa.py tests/test_gpu_primitives.py -q -m gpu > gpurun_out/r7_tests.log 2>&1
echo "tests exit $?" >> gpurun_out/r7_tests.log
for m in wl star field; do
  python tools/prof_modules.py $m 64 > gpurun_out/r7_time_$m.log 2>&1
  timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r7_launches_$m.csv python tools/prof_modules.py $m 64 > gpurun_out/r7_ncu_$m.log 2>&1
done
tail -12 gpurun_out/r7_tests.log; cat gpurun_out/r7_time_*.log
