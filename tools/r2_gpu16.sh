#!/bin/bash
# round-2 run 16: why is bench.py's pageable leg (38 ms) slower than tools/bench_pageable.py (25 ms)?  NUMA binding of the process on / off
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
O=gpurun_out
numactl -H 2>/dev/null | head -6; nvidia-smi topo -m 2>/dev/null | head -6
for b in "" 1; do BIND=$b timeout 200 python tools/bench_pageable.py 2>&1 | tail -4; done | tee $O/r16_pageable.log
BIND=1 EPID_COPY_UNBIND=1 timeout 200 python tools/bench_pageable.py 2>&1 | tail -2 | tee -a $O/r16_pageable.log
