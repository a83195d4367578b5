#!/bin/bash
# round-2 run 19: 2-GPU bench line with the final build (torchrun, one rank per GPU)
cd "$GRAFT_REPO_ROOT" || exit 1
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 10 --warmup 3 2> gpurun_out/bench_2gpu_r2o.err | tail -1 > gpurun_out/bench_2gpu_r2o.json; cut -c1-250 gpurun_out/bench_2gpu_r2o.json; tail -3 gpurun_out/bench_2gpu_r2o.err
