#!/bin/bash
# Round-end evidence run on one B200 (gpurun): GPU tests, smoke, both bench arms, ncu launch lists + full captures of the dominant
# kernels.  Outputs land in gpurun_out/ (summarised into profiles/ afterwards).
set -x
tag=${1:-r2z}
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -q -m gpu > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench_$tag.err | tail -1 > gpurun_out/bench_$tag.json; cut -c1-300 gpurun_out/bench_$tag.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref_$tag.json; cut -c1-300 gpurun_out/bench_ref_$tag.json
# launch list of the device-resident pipeline (the timed region of `value`) and full captures of the dominant kernels
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 40 --csv --log-file gpurun_out/launches_$tag.csv python tools/prof_pf.py 2 512 > gpurun_out/prof_pf_$tag.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pf_win_medians -s 1 -c 1 -o gpurun_out/prof_wmed_$tag -f python tools/prof_pf.py 1 512 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pf_stream -s 1 -c 1 -o gpurun_out/prof_stream_$tag -f python tools/prof_pf.py 1 512 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pf_win_fwxm -s 1 -c 1 -o gpurun_out/prof_wfwxm_$tag -f python tools/prof_pf.py 1 512 > /dev/null 2>&1
# VMAT: launch list + the frame-streaming kernel
python tools/prof_vmat.py 256 3 > gpurun_out/vmat_time_$tag.log 2>&1; cat gpurun_out/vmat_time_$tag.log
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/launches_vmat_$tag.csv python tools/prof_vmat.py 256 1 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_vmat_front -s 1 -c 1 -o gpurun_out/prof_vmatfront_$tag -f python tools/prof_vmat.py 256 1 > /dev/null 2>&1
for m in wl star field; do python tools/prof_modules.py $m 512 > gpurun_out/time_${m}_$tag.log 2>&1; cat gpurun_out/time_${m}_$tag.log; done
ls -la gpurun_out/*$tag*
