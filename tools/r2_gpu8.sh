#!/bin/bash
# round-2 run 8: certified-noise fast re-run + overlapped re-run (tests, mixed bench), pageable staging variants, ncu source captures of the
# dominant module kernels (k_star_wobble, k_field_profile, k_wl_bb)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 900 python -m pytest tests/test_gpu_pf.py tests/test_gpu_pf_fuzz.py -q -m gpu -x > $O/r8_tests.log 2>&1; echo "tests exit $?" >> $O/r8_tests.log
tail -5 $O/r8_tests.log
timeout 300 python tools/bench_mixed.py > $O/r8_mixed.log 2>&1; cat $O/r8_mixed.log
for v in "8 1" "8 0" "12 1" "14 1"; do set -- $v; EPID_COPY_THREADS=$1 EPID_COPY_NT=$2 timeout 200 python tools/bench_pageable.py 2>&1 | tail -3; done > $O/r8_pageable.log; cat $O/r8_pageable.log
for m in star field wl; do
  timeout 200 python tools/prof_modules.py $m 256 > $O/r8_time_$m.log 2>&1; cat $O/r8_time_$m.log
done
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_star_wobble -c 1 -o $O/r8_star -f python tools/prof_modules.py star 64 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_field_profile -c 1 -o $O/r8_fieldp -f python tools/prof_modules.py field 256 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_field_center -c 1 -o $O/r8_fieldc -f python tools/prof_modules.py field 256 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_wl_bb -c 1 -o $O/r8_wlbb -f python tools/prof_modules.py wl 256 > /dev/null 2>&1
ls -la $O
