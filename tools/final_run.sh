#!/bin/bash
# Round-end evidence run on one B200 (gpurun): GPU tests, smoke, both bench arms, ncu launch lists + full captures of the dominant
# kernels.  Outputs land in gpurun_out/ (summarised into profiles/ afterwards).
set -x
tag=${1:-r2z}
mkdir -p gpurun_out
O=gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > $O/pytest_gpu_$tag.log 2>&1; tail -3 $O/pytest_gpu_$tag.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 2> $O/bench_$tag.err | tail -1 > $O/bench_$tag.json; cut -c1-300 $O/bench_$tag.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 > $O/bench_ref_$tag.json; cut -c1-300 $O/bench_ref_$tag.json
# launch list of the device-resident pipeline (the timed region of `value`) and full captures of the dominant kernels
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 40 --csv --log-file $O/launches_$tag.csv python tools/prof_pf.py 2 512 > $O/prof_pf_$tag.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pf_win_medians -s 1 -c 1 -o $O/prof_wmed_$tag -f python tools/prof_pf.py 1 512 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pf_stream -s 1 -c 1 -o $O/prof_stream_$tag -f python tools/prof_pf.py 1 512 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_inv_stream -s 1 -c 1 -o $O/prof_inv_$tag -f python tools/prof_modules.py field 512 > /dev/null 2>&1
# modules: wall time + launch lists
for m in "wl 2048" "star 256" "field 4096"; do python tools/prof_modules.py $m > $O/time_${m%% *}_$tag.log 2>&1; cat $O/time_${m%% *}_$tag.log; done
python tools/prof_vmat.py 1024 3 > $O/vmat_time_$tag.log 2>&1; cat $O/vmat_time_$tag.log
for m in wl star field; do timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_${m}_$tag.csv python tools/prof_modules.py $m 512 > /dev/null 2>&1; done
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file $O/launches_vmat_$tag.csv python tools/prof_vmat.py 256 1 > /dev/null 2>&1
timeout 200 python tools/bench_mixed.py > $O/mixed_$tag.log 2>&1; tail -4 $O/mixed_$tag.log
ls -la $O/*$tag*
