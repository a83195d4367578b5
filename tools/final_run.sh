#!/bin/bash
# Round-end evidence run on one B200 (gpurun): GPU tests, smoke, both bench arms, ncu launch list + full captures of the dominant
# kernels.  Outputs land in gpurun_out/ (summarised into profiles/ afterwards).
set -x
tag=${1:-r2z}
timeout 1200 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 600 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench_$tag.err | tail -1 > gpurun_out/bench_$tag.json; cut -c1-300 gpurun_out/bench_$tag.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref_$tag.json; cut -c1-300 gpurun_out/bench_ref_$tag.json
# launch list of the device-resident pipeline (the timed region of `value`) and full captures of the two dominant kernels
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 40 --csv --log-file gpurun_out/launches_$tag.csv python tools/prof_pf.py 2 512 > gpurun_out/prof_pf_$tag.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pf_win_medians -s 1 -c 1 -o gpurun_out/prof_wmed_$tag -f python tools/prof_pf.py 1 512 > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pf_stream -s 1 -c 1 -o gpurun_out/prof_stream_$tag -f python tools/prof_pf.py 1 512 > /dev/null 2>&1
ls -la gpurun_out/*$tag*
