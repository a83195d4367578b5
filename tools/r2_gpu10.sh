#!/bin/bash
# round-2 run 10: full GPU suite after the certified inversion statistics / find_peaks stage 1 / WL atomics; module timings; ncu
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > $O/r10_tests.log 2>&1; echo "tests exit $?" >> $O/r10_tests.log
tail -15 $O/r10_tests.log
timeout 200 python tools/prof_pf.py 10 512 > $O/r10_pf.log 2>&1; cat $O/r10_pf.log
for m in "star 256" "field 4096" "wl 2048"; do
  timeout 300 python tools/prof_modules.py $m > $O/r10_time_${m// /_}.log 2>&1; cat $O/r10_time_${m// /_}.log
done
timeout 200 python tools/prof_vmat.py 1024 3 > $O/r10_vmat.log 2>&1; cat $O/r10_vmat.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r10_launches_star.csv python tools/prof_modules.py star 256 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r10_launches_field.csv python tools/prof_modules.py field 512 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r10_launches_wl.csv python tools/prof_modules.py wl 512 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:k_inv_stream -s 1 -c 1 -o $O/r10_inv -f python tools/prof_modules.py field 512 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k k_wl_bb -c 1 -o $O/r10_wlbb -f python tools/prof_modules.py wl 256 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k k_field_profile -c 1 -o $O/r10_fieldp -f python tools/prof_modules.py field 256 > /dev/null 2>&1
ls -la $O | tail -12
