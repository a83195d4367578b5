#!/bin/bash
# round-2 run 12: full GPU suite after the widths / radix-select / warp-cooperative Nelder-Mead changes; module timings and launch lists
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > $O/r12_tests.log 2>&1; echo "tests exit $?" >> $O/r12_tests.log
tail -15 $O/r12_tests.log
timeout 200 python tools/prof_pf.py 10 512 > $O/r12_pf.log 2>&1; cat $O/r12_pf.log
for m in "star 256" "field 4096" "wl 2048"; do
  timeout 300 python tools/prof_modules.py $m > $O/r12_time_${m// /_}.log 2>&1; cat $O/r12_time_${m// /_}.log
done
timeout 200 python tools/prof_vmat.py 1024 3 2>&1 | tee $O/r12_vmat.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r12_launches_vmat.csv python tools/prof_vmat.py 256 1 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r12_launches_field.csv python tools/prof_modules.py field 512 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r12_launches_star.csv python tools/prof_modules.py star 256 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k k_field_profile -c 1 -o $O/r12_fieldp -f python tools/prof_modules.py field 256 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k k_field_center -c 1 -o $O/r12_fieldc -f python tools/prof_modules.py field 256 > /dev/null 2>&1
ls -la $O | tail -8
