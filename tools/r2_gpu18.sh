#!/bin/bash
# round-2 run 18: bench line with the forking parts of bench.py isolated in a spawned process
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 900 python bench.py --steps 20 --warmup 5 2> gpurun_out/bench_r2o.err | tail -1 > gpurun_out/bench_r2o.json; cut -c1-200 gpurun_out/bench_r2o.json; tail -3 gpurun_out/bench_r2o.err
