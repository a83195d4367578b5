#!/bin/bash
# round-2 run 15: PF translation units without the find_peaks skip table; refreshed bench line (value, e2e, e2e_pageable, mixed, modules)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
tag=r2n
timeout 900 python -m pytest tests/test_gpu_pf.py tests/test_gpu_pf_fuzz.py tests/test_gpu_primitives.py -q -m gpu > $O/pytest_pf_$tag.log 2>&1; echo "tests exit $?" >> $O/pytest_pf_$tag.log; tail -3 $O/pytest_pf_$tag.log
timeout 120 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
timeout 900 python bench.py --steps 20 --warmup 5 2> $O/bench_$tag.err | tail -1 > $O/bench_$tag.json; cut -c1-200 $O/bench_$tag.json
timeout 300 python bench.py --impl reference --steps 5 --warmup 1 2>/dev/null | tail -1 > $O/bench_ref_$tag.json; cut -c1-200 $O/bench_ref_$tag.json
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 12 -c 40 --csv --log-file $O/launches_$tag.csv python tools/prof_pf.py 2 512 > $O/prof_pf_$tag.log 2>&1; cat $O/prof_pf_$tag.log
timeout 300 ncu --set full --clock-control none --import-source on -k k_star_rows -s 1 -c 1 -o $O/prof_starrows_$tag -f python tools/prof_modules.py star 256 > /dev/null 2>&1
timeout 200 python tools/bench_mixed.py > $O/mixed_$tag.log 2>&1; tail -3 $O/mixed_$tag.log | cut -c1-120
