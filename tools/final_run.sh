#!/bin/bash
# Round-end evidence run on one B200 (gpurun): GPU tests, smoke, both bench arms, ncu launch list + full captures.
# Outputs land in gpurun_out/ (copied into profiles/ afterwards).
set -x
tag=${1:-r1m}
timeout 900 python -m pytest tests -q -m gpu -x > gpurun_out/pytest_gpu_$tag.log 2>&1; tail -3 gpurun_out/pytest_gpu_$tag.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 400 python bench.py --steps 10 --warmup 3 2> gpurun_out/bench_$tag.err | tail -1 > gpurun_out/bench_$tag.json; cut -c1-300 gpurun_out/bench_$tag.json
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 2>/dev/null | tail -1 > gpurun_out/bench_ref_$tag.json; cut -c1-200 gpurun_out/bench_ref_$tag.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file gpurun_out/launches_$tag.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline > gpurun_out/bench_ncu_$tag.log 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pf_stream -s 3 -c 1 -o gpurun_out/prof_stream_$tag python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
timeout 300 ncu --set full --clock-control none --import-source on -k regex:k_pf_windows_fast -s 3 -c 1 -o gpurun_out/prof_windows_$tag python bench.py --steps 1 --warmup 3 --no-cpu-baseline > /dev/null 2>&1
ls -la gpurun_out/*$tag*
