#!/bin/bash
# round-2 run 11: full GPU suite after the parallel distance stage (find_peaks) and the run-based labelling (k_wl_bb); VMAT A/B against the
# r2k library; Starshot 2 vs 3 CTAs per SM; module timings and launch lists
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
timeout 1500 python -m pytest tests -q -m gpu > $O/r11_tests.log 2>&1; echo "tests exit $?" >> $O/r11_tests.log
tail -15 $O/r11_tests.log
timeout 200 python tools/prof_pf.py 10 512 > $O/r11_pf.log 2>&1; cat $O/r11_pf.log
for m in "star 256" "field 4096" "wl 2048"; do
  timeout 300 python tools/prof_modules.py $m > $O/r11_time_${m// /_}.log 2>&1; cat $O/r11_time_${m// /_}.log
done
EPID_LIB=$PWD/variants/libepid_star3.so timeout 300 python tools/prof_modules.py star 256 2>&1 | head -1 | sed 's/^/star3: /' | tee $O/r11_star3.log
timeout 200 python tools/prof_vmat.py 1024 3 2>&1 | tee $O/r11_vmat.log
EPID_LIB=$PWD/variants/libepid_r2k.so timeout 200 python tools/prof_vmat.py 1024 3 2>&1 | sed 's/^/r2k lib: /' | tee -a $O/r11_vmat.log
EPID_LIB=$PWD/variants/libepid_r2k.so timeout 200 python tools/prof_modules.py field 512 2>&1 | sed 's/^/r2k lib: /' | tee -a $O/r11_vmat.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r11_launches_vmat.csv python tools/prof_vmat.py 256 1 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r11_launches_field.csv python tools/prof_modules.py field 512 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r11_launches_wl.csv python tools/prof_modules.py wl 512 > /dev/null 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/r11_launches_star.csv python tools/prof_modules.py star 256 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k k_wl_bb -c 1 -o $O/r11_wlbb -f python tools/prof_modules.py wl 256 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k k_field_profile -c 1 -o $O/r11_fieldp -f python tools/prof_modules.py field 256 > /dev/null 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k k_vmat_profile -c 1 -o $O/r11_vmatp -f python tools/prof_vmat.py 256 1 > /dev/null 2>&1
ls -la $O | tail -14
