#!/bin/bash
# round-2 run 9: full GPU suite after the find_peaks / Starshot / re-run changes, module timings, ncu source captures
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > $O/r9_tests.log 2>&1; echo "tests exit $?" >> $O/r9_tests.log
tail -15 $O/r9_tests.log
timeout 200 python tools/prof_pf.py 10 512 > $O/r9_pf.log 2>&1; cat $O/r9_pf.log
for m in "star 256" "field 512" "field 4096" "wl 2048"; do
  timeout 300 python tools/prof_modules.py $m > $O/r9_time_${m// /_}.log 2>&1; cat $O/r9_time_${m// /_}.log
done
timeout 200 python tools/prof_vmat.py 256 3 > $O/r9_vmat.log 2>&1; cat $O/r9_vmat.log
timeout 400 ncu --set full --clock-control none --import-source on -k k_star_rows -s 1 -c 1 -o $O/r9_star -f python tools/prof_modules.py star 64 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k k_field_profile -c 1 -o $O/r9_fieldp -f python tools/prof_modules.py field 256 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k k_wl_bb -c 1 -o $O/r9_wlbb -f python tools/prof_modules.py wl 256 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r9_launches_star.csv python tools/prof_modules.py star 256 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r9_launches_field.csv python tools/prof_modules.py field 512 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r9_launches_wl.csv python tools/prof_modules.py wl 512 > /dev/null 2>&1
ls -la $O | tail -20
