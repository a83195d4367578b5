#!/bin/bash
# round-2 run 13: Starshot ring sampling with independent loads; 2 vs 3 resident CTAs for k_star_rows / k_field_profile
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 600 python -m pytest tests/test_gpu_starshot.py tests/test_gpu_field.py tests/test_gpu_primitives.py -q -m gpu > $O/r13_tests.log 2>&1; echo "tests exit $?" >> $O/r13_tests.log
tail -4 $O/r13_tests.log
for lib in "" "$PWD/variants/libepid_occ3.so"; do
  for m in "star 256" "field 4096"; do
    EPID_LIB=$lib timeout 300 python tools/prof_modules.py $m 2>&1 | head -1 | sed "s#^#lib=${lib##*/} #" | tee -a $O/r13_times.log
  done
done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r13_launches_star.csv python tools/prof_modules.py star 256 > /dev/null 2>&1
