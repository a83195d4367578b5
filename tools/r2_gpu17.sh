#!/bin/bash
# round-2 run 17: bench.py's pageable leg with and without the CPU baseline (128 forked workers) in the same process
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
for f in "--no-cpu-baseline" ""; do
  timeout 400 python bench.py --steps 10 --warmup 3 --no-modules $f 2>/dev/null | tail -1 | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('flags=[$f]', 'value', round(d['value']), 'e2e ms', round(d['e2e']['ms_per_step'],2), 'pageable ms', round(d['e2e_pageable']['ms_per_step'],2), 'frac', round(d['e2e_pageable']['frac_of_pinned'],3))" | tee -a $O/r17_bench_pageable.log
done
